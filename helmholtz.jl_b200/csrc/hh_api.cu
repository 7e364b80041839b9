// hh_api.cu -- the extern "C" boundary of libhelmholtz_b200.so (see include/helmholtz_b200.h).
// Catches every C++ exception, maps it to a status code + message, shards right-hand sides over
// devices for multi-GPU handles, and implements the host-side set-up helpers of the reference
// (getABL, getMaximalFrequency, loc2cs) that stay Float64 host work.
#include <array>
#include <memory>
#include <mutex>
#include <thread>

#include "hh_solver.cuh"

using namespace hh;

// Device-side staging of host right-hand sides / solutions for hh_solve: two slots so that the copies of one
// sub-batch overlap the solve of the other (copy stream + events), kept across calls.
struct HostStage {
    cudaStream_t cs = nullptr;
    cudaEvent_t h2d[2] = {nullptr, nullptr}, solved[2] = {nullptr, nullptr}, d2h[2] = {nullptr, nullptr};
    DevBuf<char> db[2], dx[2];
    void init() {
        if (cs) return;
        HH_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));  // must not serialise with the legacy stream
        for (int i = 0; i < 2; ++i) {
            HH_CUDA(cudaEventCreateWithFlags(&h2d[i], cudaEventDisableTiming));
            HH_CUDA(cudaEventCreateWithFlags(&solved[i], cudaEventDisableTiming));
            HH_CUDA(cudaEventCreateWithFlags(&d2h[i], cudaEventDisableTiming));
        }
    }
    void ensure(size_t bytes) {
        for (int i = 0; i < 2; ++i)
            if (db[i].n < bytes) {
                db[i].alloc(bytes);
                dx[i].alloc(bytes);
            }
    }
    void release() {
        for (int i = 0; i < 2; ++i) {
            db[i].release();
            dx[i].release();
        }
    }
    ~HostStage() {
        for (int i = 0; i < 2; ++i) {
            if (h2d[i]) cudaEventDestroy(h2d[i]);
            if (solved[i]) cudaEventDestroy(solved[i]);
            if (d2h[i]) cudaEventDestroy(d2h[i]);
        }
        if (cs) cudaStreamDestroy(cs);
    }
};

struct hh_handle_s {
    std::vector<std::unique_ptr<SolverBase>> subs;  // one per device
    std::vector<std::unique_ptr<HostStage>> stages; // one per device
    // HH_C64_MIXED: the ComplexF32 companion of every replica (owns the hierarchy) and its staging blocks
    std::vector<std::unique_ptr<SolverBase>> lows;
    std::vector<std::unique_ptr<DevBuf<cx<float>>>> lo_b, lo_z;
    int precision = HH_C64;
    Problem pb;  // the whole grid (also for slab handles, whose replicas each describe one slab)
    // slab decomposition of one problem (hh_slab.cuh): 0 = none (subs are replicas that share the right-hand sides out),
    // 1 = one host thread per slab inside this process (caller arrays are whole-grid arrays),
    // 2 = this process holds one slab, NCCL between the processes (caller arrays hold the owned planes)
    int slab_mode = 0;
    std::shared_ptr<ThreadGroup> grp;
    int64_t model_plane0 = 0, model_planes = 0;  // planes of the last dimension the caller's m / gamma arrays hold
    std::string err;
    bool have_opts = false;
    hh_mg_options opts{};
};

static thread_local std::string g_err;

template <class F>
static int guarded(hh_handle_t h, F&& f) {
    try {
        return f();
    } catch (const hh::Error& e) {
        g_err = e.what();
        if (h) h->err = g_err;
        return e.code;
    } catch (const std::bad_alloc&) {
        g_err = "host allocation failed";
        if (h) h->err = g_err;
        return HH_ERR_ALLOC;
    } catch (const std::exception& e) {
        g_err = e.what();
        if (h) h->err = g_err;
        return HH_ERR_STATE;
    } catch (...) {
        g_err = "unknown error";
        if (h) h->err = g_err;
        return HH_ERR_STATE;
    }
}

// run f(sub_index) on every device replica concurrently; rethrow the first failure
template <class F>
static void for_each_sub(hh_handle_t h, F&& f) {
    const int n = (int)h->subs.size();
    if (n == 1) {
        f(0);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::string> msg(n);
    std::vector<int> code(n, 0);
    for (int i = 0; i < n; ++i) {
        th.emplace_back([&, i] {
            try {
                f(i);
            } catch (const hh::Error& e) {
                code[i] = e.code;
                msg[i] = e.what();
            } catch (const std::exception& e) {
                code[i] = HH_ERR_STATE;
                msg[i] = e.what();
            }
            if (code[i] != 0 && h->grp) h->grp->fail();  // slabs run in lockstep: release the ones waiting on this one
        });
    }
    for (auto& t : th) t.join();
    if (h->grp) h->grp->reset();
    // report the slab that failed first-hand rather than the ones it released
    for (int i = 0; i < n; ++i)
        if (code[i] != 0 && msg[i].find("another slab") == std::string::npos)
            throw hh::Error(code[i], "device replica " + std::to_string(i) + ": " + msg[i]);
    for (int i = 0; i < n; ++i)
        if (code[i] != 0) throw hh::Error(code[i], "device replica " + std::to_string(i) + ": " + msg[i]);
}

extern "C" {

int hh_version(void) { return HH_VERSION; }

const char* hh_last_error(hh_handle_t h) { return h ? h->err.c_str() : g_err.c_str(); }

int hh_device_count(int* count) {
    if (!count) return HH_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return HH_OK;
}

// ---------------------------------------------------------------------------------------------
// getABL (src/GetHelmholtz.jl:97-220), live branches only: 2-D impl == 1 (:141-163), 3-D (:164-218)
// ---------------------------------------------------------------------------------------------
int hh_get_abl(int dim, const int64_t* n, int neumann_on_top, const int64_t* pad, double amp, double* gamma) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE((dim == 2 || dim == 3) && n && pad && gamma, HH_ERR_ARG, "hh_get_abl: bad arguments");
        std::vector<double> tab[4];
        abl_tables(dim, n, neumann_on_top, pad, tab);  // the 1-D ramps (shared with the device-side hh_set_frequency_abl)
        if (dim == 2) {
            // the reference adds the side ramps over all columns and subtracts the corner products only where
            // the side ramp meets the dim-2 ramp (first p1 / last p1 rows separately, GetHelmholtz.jl:152-161)
            const std::vector<double>&l1 = tab[0], &r1 = tab[1], &top = tab[2], &bot = tab[3];
            for (int64_t j = 0; j < n[1]; ++j)
                for (int64_t i = 0; i < n[0]; ++i) {
                    double g = top[j] - l1[i] * top[j] - r1[i] * top[j];
                    g += bot[j];
                    g += l1[i];
                    g += r1[i];
                    g -= l1[i] * bot[j];
                    g -= r1[i] * bot[j];
                    gamma[i + n[0] * j] = g * amp;
                }
            return HH_OK;
        }
        for (int64_t k = 0; k < n[2]; ++k)
            for (int64_t j = 0; j < n[1]; ++j)
                for (int64_t i = 0; i < n[0]; ++i) {
                    double v = (tab[0][i] + tab[1][j] + tab[2][k]) * amp;
                    if (v >= amp) v = amp;
                    gamma[i + n[0] * (j + n[1] * k)] = v;
                }
        return HH_OK;
    });
}

// getMaximalFrequency (src/GetHelmholtz.jl:75-79)
int hh_get_maximal_frequency(const double* m, int64_t n, int dim, const double* h, double* omega_max) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE(m && h && omega_max && n > 0 && dim >= 1 && dim <= 3, HH_ERR_ARG, "hh_get_maximal_frequency: bad arguments");
        double mm = m[0], hm = h[0];
        for (int64_t i = 1; i < n; ++i) mm = std::max(mm, m[i]);
        for (int d = 1; d < dim; ++d) hm = std::max(hm, h[d]);
        *omega_max = (0.1 * 2 * M_PI) / (hm * std::sqrt(mm));
        return HH_OK;
    });
}

// loc2cs / loc2cs3D (src/getPointSource.jl:82-102): 1-based in, 1-based out
int64_t hh_point_source_index(int dim, const int64_t* n, const int64_t* sub) {
    if (!n || !sub) return -1;
    if (dim == 2) return sub[0] + (sub[1] - 1) * n[0];
    if (dim == 3) return sub[0] + (sub[1] - 1) * n[0] + (sub[2] - 1) * n[0] * n[1];
    return -1;
}

// one solver (plus its ComplexF32 companion when the precision is mixed) for the grid `pb` on `device`
static void add_replica(hh_handle_s* h, const Problem& pb, int precision, int device) {
    const size_t i = h->subs.size();
    if (precision == HH_C32) h->subs.emplace_back(new Solver<float>(pb, device));
    else h->subs.emplace_back(new Solver<double>(pb, device));
    h->stages.emplace_back(new HostStage);
    if (precision == HH_C64_MIXED) {
        h->lows.emplace_back(new Solver<float>(pb, device));
        h->lo_b.emplace_back(new DevBuf<cx<float>>);
        h->lo_z.emplace_back(new DevBuf<cx<float>>);
        SolverBase* hi = h->subs[i].get();
        SolverBase* lo = h->lows[i].get();
        DevBuf<cx<float>>* bb = h->lo_b[i].get();
        DevBuf<cx<float>>* zz = h->lo_z[i].get();
        hi->krylov_only = true;
        // z = M(b): b, z ComplexF64 blocks of the outer solver; the cycle runs on ComplexF32 copies
        hi->prec_hook = [hi, lo, bb, zz](const void* b, void* z, int nrhs) {
            lo->stream = hi->stream;
            lo->ensure_cycle_memory(nrhs);
            const int64_t ldl = lo->internal_ld(), ldh = hi->internal_ld();
            const size_t need = (size_t)ldl * nrhs;
            if (bb->n < need) {
                bb->alloc(need);
                zz->alloc(need);
                HH_CUDA(cudaMemsetAsync(bb->p, 0, need * sizeof(cx<float>), hi->stream));  // ghost nodes stay zero
                HH_CUDA(cudaMemsetAsync(zz->p, 0, need * sizeof(cx<float>), hi->stream));
            }
            const int n0 = hi->pb.n[0];
            const int64_t rows = (int64_t)hi->pb.n[1] * hi->pb.n[2];
            dim3 g(std::max(1, 592 / std::max(nrhs, 1)), nrhs);
            k_convert<double, float><<<g, 256, 0, hi->stream>>>((const cx<double>*)b, bb->p, n0, rows, hi->internal_pitch(),
                                                               lo->internal_pitch(), ldh, ldl);
            lo->precondition_internal(bb->p, zz->p, nrhs);
            k_convert<float, double><<<g, 256, 0, hi->stream>>>(zz->p, (cx<double>*)z, n0, rows, lo->internal_pitch(),
                                                               hi->internal_pitch(), ldl, ldh);
            hi->launches += 2;
        };
    }
}

// ---------------------------------------------------------------------------------------------
// GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72) on the spread nodal Laplacian and mass of
// src/PlainNodalLaplacian.jl:49-141, as a stored 3^dim-point stencil.  The Kronecker construction
//   Lap = G' Gs,  Gs = (1-b) [spread gradients] + b G,   H = Lap + M Diagonal(mass)
// collapses to sums of tensor products of 1-D tridiagonal matrices: T_d = ddx' ddx (the first-order Neumann
// Laplacian), A_d = av3term(n_d+1, 1/2) (spreading), B_d = av3term(n_d+1, beta_mass):
//   2-D  Lap = T1 (x) [(1-b) A2 + b I] + [(1-b) A1 + b I] (x) T2,            M = 1/2 (B2 (x) I + I (x) B1)
//   3-D  Lap = sum_d T_d (x) [b I(x)I + (1-b)/2 (A_e (x) I + I (x) A_f)],     M = 1/3 sum_d B_d (x) I (x) I
// Host-side Float64 set-up like getABL; no device is needed.
// ---------------------------------------------------------------------------------------------
extern "C" int hh_ho_stencil(int dim, const int64_t* n_nodes, const double* hsp, const double* m, const double* gamma,
                             double wre, double wim, int neumann_on_top, int sommerfeld, const double* beta, double* coef_out) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE(n_nodes && hsp && m && gamma && beta && coef_out, HH_ERR_ARG, "hh_ho_stencil: bad arguments");
        build_ho_stencil(dim, n_nodes, hsp, m, gamma, wre, wim, neumann_on_top, sommerfeld, beta, coef_out);
        return HH_OK;
    });
}

// GetHelmholtzOperator / GetHelmholtzOperatorHO as the sparse matrix the reference returns (SparseMatrixCSC, 0-based
// here): call with rowval = nzval = NULL to get colptr and nnz = colptr[N], then again with arrays of that size.
extern "C" int hh_assemble_csc(int dim, const int64_t* n_nodes, const double* hsp, const double* m, const double* gamma,
                               double wre, double wim, int neumann_on_top, int sommerfeld, int order_bc, double shift,
                               const double* ho_beta, int64_t* colptr, int64_t* rowval, double* nzval) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE(n_nodes && hsp && m && gamma && colptr && ((rowval == nullptr) == (nzval == nullptr)), HH_ERR_ARG,
                   "hh_assemble_csc: bad arguments");
        HH_REQUIRE(dim == 2 || dim == 3, HH_ERR_ARG, "hh_assemble_csc: dim must be 2 or 3");
        int64_t N = 1;
        for (int d = 0; d < dim; ++d) N *= n_nodes[d];
        const int NS = dim == 3 ? 27 : 9;
        std::vector<double> coef((size_t)2 * NS * N);
        if (ho_beta) {
            build_ho_stencil(dim, n_nodes, hsp, m, gamma, wre, wim, neumann_on_top, sommerfeld, ho_beta, coef.data());
            const int center = dim == 3 ? 13 : 4;
            for (int64_t p = 0; p < N; ++p) coef[2 * ((int64_t)center * N + p) + 1] += shift * wre * wre * m[p];
        } else {
            build_plain_stencil(dim, n_nodes, hsp, m, gamma, wre, wim, neumann_on_top, sommerfeld, order_bc, shift, coef.data());
        }
        stencil_to_csc(dim, n_nodes, coef.data(), colptr, rowval, nzval);
        return HH_OK;
    });
}

extern "C" int hh_stencil_adjoint(int dim, const int64_t* n_nodes, const double* coef_in, double* coef_out) {
    return guarded(nullptr, [&]() -> int {
        adjoint_stencil(dim, n_nodes, coef_in, coef_out);
        return HH_OK;
    });
}

// ---------------------------------------------------------------------------------------------
// slab_mode 0: `ndev` replicas of the whole grid.  1: `ndev` slabs of one grid inside this process.  2: slab `rank` of
// `nranks` on devices[0], NCCL communicator from `uid`.
static int create_impl(int dim, const int64_t* n_nodes, const double* hsp, const double* m, const double* gamma,
                       double wre, double wim, int neumann_on_top, int sommerfeld, int order_bc, int precision,
                       const int* devices, int ndev, hh_handle_t* out, int slab_mode = 0, int levels = 0, int rank = 0,
                       int nranks = 1, const void* uid = nullptr, int64_t model_plane0 = 0, int64_t model_planes = 0) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE(out != nullptr, HH_ERR_ARG, "hh_create: out is NULL");
        *out = nullptr;
        HH_REQUIRE(dim == 2 || dim == 3, HH_ERR_ARG, "hh_create: dim must be 2 or 3");
        HH_REQUIRE(n_nodes && hsp && m && gamma, HH_ERR_ARG, "hh_create: NULL array");
        HH_REQUIRE(precision == HH_C64 || precision == HH_C32 || precision == HH_C64_MIXED, HH_ERR_ARG, "hh_create: bad precision");
        HH_REQUIRE(order_bc == 1 || order_bc == 2, HH_ERR_ARG, "getNodalLaplacianMatrix: BC not supported");
        HH_REQUIRE(wre != 0.0, HH_ERR_ARG, "hh_create: Re(omega) must be non-zero");
        HH_REQUIRE(ndev >= 1 && devices, HH_ERR_ARG, "hh_create: no devices");
        Problem pb;
        pb.dim = dim;
        for (int d = 0; d < dim; ++d) {
            HH_REQUIRE(n_nodes[d] >= 2 && n_nodes[d] < (1 << 30), HH_ERR_ARG, "hh_create: node counts must be >= 2");
            HH_REQUIRE(hsp[d] > 0.0, HH_ERR_ARG, "hh_create: mesh spacing must be positive");
            pb.n[d] = (int)n_nodes[d];
            pb.h[d] = hsp[d];
        }
        pb.neumann_top = neumann_on_top ? 1 : 0;
        pb.sommerfeld = sommerfeld ? 1 : 0;
        pb.order_bc = order_bc;
        int ndevices = 0;
        cudaError_t e = cudaGetDeviceCount(&ndevices);
        if (e != cudaSuccess || ndevices == 0) {
            cudaGetLastError();
            throw hh::Error(HH_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
        }
        std::unique_ptr<hh_handle_s> h(new hh_handle_s);
        h->precision = precision;
        h->pb = pb;
        h->slab_mode = slab_mode;
        for (int i = 0; i < ndev; ++i)
            HH_REQUIRE(devices[i] >= 0 && devices[i] < ndevices, HH_ERR_ARG, "hh_create: bad device ordinal");
        if (slab_mode == 0) {
            for (int i = 0; i < ndev; ++i) add_replica(h.get(), pb, precision, devices[i]);
        } else {
            HH_REQUIRE(dim == 3, HH_ERR_UNSUPPORTED, "slab decomposition needs a 3-D grid");
            HH_REQUIRE(levels >= 1 && levels <= HH_MAX_LEVELS, HH_ERR_ARG, "slab decomposition: levels out of range");
            const int nr = slab_mode == 1 ? ndev : nranks;
            HH_REQUIRE(slab_mode == 1 || (ndev == 1 && uid != nullptr && rank >= 0 && rank < nranks), HH_ERR_ARG,
                       "hh_create_slab_nccl: bad rank / communicator id");
            if (slab_mode == 1) h->grp = std::make_shared<ThreadGroup>(nr);
            for (int q = 0; q < (slab_mode == 1 ? ndev : 1); ++q) {
                const int r = slab_mode == 1 ? q : rank;
                std::vector<SlabLevel> geo;
                HH_REQUIRE(slab_partition(pb.n[2], levels, nr, r, geo), HH_ERR_ARG,
                           "slab decomposition: the cells of the last dimension must be divisible by 2^(levels-1) and the "
                           "coarsest level needs at least one cell per slab");
                Problem pl = pb;
                pl.n[2] = geo[0].nloc;
                add_replica(h.get(), pl, precision, devices[q]);
                std::shared_ptr<SlabTransport> tr;
                if (slab_mode == 1) {
                    tr = std::make_shared<ThreadTransport>(r, h->grp);
                } else {
                    HH_CUDA(cudaSetDevice(devices[q]));
                    tr = std::make_shared<NcclTransport>(r, nr, uid);
                }
                h->subs[q]->slab = tr;
                h->subs[q]->sgeo = geo;
                if (!h->lows.empty()) {
                    h->lows[q]->slab = tr;
                    h->lows[q]->sgeo = geo;
                }
            }
        }
        h->model_plane0 = model_plane0;
        h->model_planes = model_planes > 0 ? model_planes : pb.n[2];
        if (slab_mode)
            for (auto& sb : h->subs)
                HH_REQUIRE(sb->sgeo[0].koff >= h->model_plane0 && sb->sgeo[0].koff + sb->sgeo[0].nloc <= h->model_plane0 + h->model_planes,
                           HH_ERR_ARG, "slab decomposition: m / gamma do not hold all planes of the slab (halo planes "
                                       "included; see hh_slab_partition: koff <= k < koff + nloc)");
        for_each_sub(h.get(), [&](int i) {
            // a slab reads its planes (halo planes included) out of the caller's arrays
            const int64_t off = slab_mode ? ((int64_t)h->subs[i]->sgeo[0].koff - h->model_plane0) * pb.n[0] * pb.n[1] : 0;
            h->subs[i]->set_model(m + off, gamma + off, wre, wim);
            if (!h->lows.empty()) h->lows[i]->set_model(m + off, gamma + off, wre, wim);
        });
        h->pb.w_re = wre;
        h->pb.w_im = wim;
        *out = h.release();
        return HH_OK;
    });
}

int hh_create(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma, double omega_re,
              double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc, int precision, int device,
              hh_handle_t* out) {
    return create_impl(dim, n_nodes, h, m, gamma, omega_re, omega_im, neumann_on_top, sommerfeld, order_neumann_bc,
                       precision, &device, 1, out);
}

int hh_create_multi(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                    double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                    int precision, const int* devices, int n_devices, hh_handle_t* out) {
    return create_impl(dim, n_nodes, h, m, gamma, omega_re, omega_im, neumann_on_top, sommerfeld, order_neumann_bc,
                       precision, devices, n_devices, out);
}

int hh_slab_partition(int64_t n3_nodes, int levels, int nranks, int rank, int64_t* out) {
    if (!out || levels < 1 || levels > HH_MAX_LEVELS || n3_nodes >= (1 << 30)) return HH_ERR_ARG;
    std::vector<SlabLevel> geo;
    if (!slab_partition((int)n3_nodes, levels, nranks, rank, geo)) return HH_ERR_ARG;
    for (int l = 0; l < levels; ++l) {
        const SlabLevel& g = geo[l];
        const int64_t v[7] = {g.own0, g.own1, g.koff, g.nloc, g.zb, g.ze, g.n2g};
        for (int q = 0; q < 7; ++q) out[7 * l + q] = v[q];
    }
    return HH_OK;
}

int hh_nccl_unique_id(void* id128) {
    return guarded(nullptr, [&]() -> int {
        HH_REQUIRE(id128 != nullptr, HH_ERR_ARG, "hh_nccl_unique_id: NULL");
        NcclApi& api = NcclApi::get();
        HH_REQUIRE(api.ok, HH_ERR_UNSUPPORTED, api.error);
        ncclUniqueId id;
        HH_NCCL(api.GetUniqueId(&id));
        memcpy(id128, &id, sizeof(id));
        return HH_OK;
    });
}

int hh_create_slab_local(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                         double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                         int precision, const int* devices, int n_slabs, int levels, hh_handle_t* out) {
    return create_impl(dim, n_nodes, h, m, gamma, omega_re, omega_im, neumann_on_top, sommerfeld, order_neumann_bc,
                       precision, devices, n_slabs, out, 1, levels);
}

int hh_create_slab_nccl(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                        double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                        int precision, int device, int levels, int rank, int nranks, const void* unique_id,
                        int64_t model_plane0, int64_t model_planes, hh_handle_t* out) {
    return create_impl(dim, n_nodes, h, m, gamma, omega_re, omega_im, neumann_on_top, sommerfeld, order_neumann_bc,
                       precision, &device, 1, out, 2, levels, rank, nranks, unique_id, model_plane0, model_planes);
}

int hh_slab_info(hh_handle_t h, int* mode, int* n_slabs, int* rank, int64_t* own0, int64_t* own1) {
    if (!h) return HH_ERR_ARG;
    const SolverBase* s = h->subs[0].get();
    if (mode) *mode = h->slab_mode;
    if (n_slabs) *n_slabs = s->slab ? s->slab->nranks : 1;
    if (rank) *rank = (h->slab_mode == 2) ? s->slab->rank : 0;
    // planes of the last dimension the caller's arrays hold
    if (own0) *own0 = (h->slab_mode == 2) ? s->sgeo[0].own0 : 0;
    if (own1) *own1 = (h->slab_mode == 2) ? s->sgeo[0].own1 : h->pb.n[2];
    return HH_OK;
}

int hh_destroy(hh_handle_t h) {
    if (!h) return HH_OK;
    return guarded(nullptr, [&]() -> int {
        for (size_t i = 0; i < h->subs.size(); ++i) {
            cudaSetDevice(h->subs[i]->device);
            h->stages[i].reset();
            if (!h->lows.empty()) {
                h->lo_b[i].reset();
                h->lo_z[i].reset();
                h->lows[i].reset();
            }
            h->subs[i].reset();
        }
        delete h;
        return HH_OK;
    });
}

int hh_set_stream(hh_handle_t h, void* cuda_stream) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(h->subs.size() == 1 || cuda_stream == nullptr, HH_ERR_ARG,
                   "hh_set_stream: a multi-device handle runs on per-device default streams");
        h->subs[0]->stream = (cudaStream_t)cuda_stream;
        return HH_OK;
    });
}

int hh_update_model(hh_handle_t h, const double* m, const double* gamma, double omega_re, double omega_im) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(m && gamma && omega_re != 0.0, HH_ERR_ARG, "hh_update_model: bad arguments");
        for_each_sub(h, [&](int i) {
            const int64_t off = h->slab_mode ? ((int64_t)h->subs[i]->sgeo[0].koff - h->model_plane0) * h->pb.n[0] * h->pb.n[1] : 0;
            h->subs[i]->set_model(m + off, gamma + off, omega_re, omega_im);
            if (!h->lows.empty()) h->lows[i]->set_model(m + off, gamma + off, omega_re, omega_im);
            for (SolverBase* sb : {h->subs[i].get(), h->lows.empty() ? (SolverBase*)nullptr : h->lows[i].get()})
                if (sb && sb->ho) {  // the high-order stencil is rebuilt from these at the next hh_setup
                    sb->ho_m.assign(m, m + h->pb.N());
                    sb->ho_g.assign(gamma, gamma + h->pb.N());
                }
        });
        h->pb.w_re = omega_re;
        h->pb.w_im = omega_im;
        return HH_OK;
    });
}

// Frequency sweep on a resident model: omega replaced and gamma <- gamma_const + getABL(n, NeumannOnTop, pad, amp)
// evaluated on the device from the 1-D ramps (GetHelmholtz.jl:22-31 with :97-220); m is kept.  Invalidates the hierarchy.
int hh_set_frequency_abl(hh_handle_t h, double omega_re, double omega_im, double gamma_const, const int64_t* pad, double amp) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(pad && omega_re != 0.0, HH_ERR_ARG, "hh_set_frequency_abl: bad arguments");
        int64_t ng[3] = {h->pb.n[0], h->pb.n[1], h->pb.n[2]};
        for_each_sub(h, [&](int i) {
            h->subs[i]->set_frequency_abl(omega_re, omega_im, gamma_const, ng, pad, amp);
            if (!h->lows.empty()) h->lows[i]->set_frequency_abl(omega_re, omega_im, gamma_const, ng, pad, amp);
        });
        h->pb.w_re = omega_re;
        h->pb.w_im = omega_im;
        return HH_OK;
    });
}

// gamma as the device holds it (Float64 copy): whole grid, or the planes own0 <= k < own1 of an NCCL slab handle
int hh_get_gamma(hh_handle_t h, double* gamma_out) {
    if (!h || !gamma_out) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        if (h->slab_mode == 0) {
            h->subs[0]->get_gamma(gamma_out);
            return HH_OK;
        }
        const int64_t plane = (int64_t)h->pb.n[0] * h->pb.n[1];
        for (auto& sb : h->subs) {
            const SlabLevel& g = sb->sgeo[0];
            std::vector<double> loc((size_t)plane * g.nloc);
            sb->get_gamma(loc.data());
            const int64_t dst0 = h->slab_mode == 1 ? g.own0 : 0;
            std::memcpy(gamma_out + plane * dst0, loc.data() + plane * g.zb, sizeof(double) * plane * (g.own1 - g.own0));
        }
        return HH_OK;
    });
}

// getMaximalFrequency (src/GetHelmholtz.jl:75-79) from the model the handle holds on the device.  With NCCL slabs the
// maximum is over this rank's planes (halo planes included): all-reduce it with MAX over the ranks.
int hh_get_maximal_frequency_device(hh_handle_t h, double* omega_max) {
    if (!h || !omega_max) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        double mm = 0.0;
        for (auto& sb : h->subs) mm = std::max(mm, sb->max_m());
        double hm = h->pb.h[0];
        for (int d = 1; d < h->pb.dim; ++d) hm = std::max(hm, h->pb.h[d]);
        *omega_max = (0.1 * 2 * M_PI) / (hm * std::sqrt(mm));
        return HH_OK;
    });
}

// GetHelmholtzOperatorHO as the operator of this handle (enable != 0) or back to the plain operator.  m and gamma are
// the arrays hh_create was given (the library keeps Float64 host copies for the stencil construction at hh_setup).
int hh_set_operator_ho(hh_handle_t h, int enable, const double* m, const double* gamma, const double* beta) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(h->slab_mode == 0, HH_ERR_UNSUPPORTED, "the high-order operator is not available on a slab handle");
        HH_REQUIRE(!enable || (m && gamma && beta), HH_ERR_ARG, "hh_set_operator_ho: NULL array");
        const int64_t N = h->pb.N();
        for (auto* v : {&h->subs, &h->lows})
            for (auto& sb : *v) {
                HH_CUDA(cudaSetDevice(sb->device));
                sb->clear();
                sb->ho = enable != 0;
                if (enable) {
                    sb->ho_m.assign(m, m + N);
                    sb->ho_g.assign(gamma, gamma + N);
                    sb->ho_beta[0] = beta[0];
                    sb->ho_beta[1] = h->pb.dim == 3 ? beta[1] : beta[0];
                } else {
                    sb->ho_m.clear();
                    sb->ho_g.clear();
                }
            }
        return HH_OK;
    });
}

int hh_setup(hh_handle_t h, const hh_mg_options* opts) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(opts != nullptr, HH_ERR_ARG, "hh_setup: opts is NULL");
        for_each_sub(h, [&](int i) {
            h->subs[i]->setup(*opts);
            if (!h->lows.empty()) {
                h->lows[i]->stream = h->subs[i]->stream;
                h->lows[i]->setup(*opts);
            }
        });
        h->opts = *opts;
        h->have_opts = true;
        return HH_OK;
    });
}

int hh_clear(hh_handle_t h) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        for_each_sub(h, [&](int i) {
            cudaSetDevice(h->subs[i]->device);
            h->subs[i]->clear();
            h->stages[i]->release();
            if (!h->lows.empty()) {
                h->lows[i]->clear();
                h->lo_b[i]->release();
                h->lo_z[i]->release();
            }
        });
        return HH_OK;
    });
}

int hh_hierarchy_exists(hh_handle_t h) {
    if (!h) return 0;
    return (h->lows.empty() ? h->subs[0]->hierarchy_exists() : h->lows[0]->hierarchy_exists()) ? 1 : 0;
}

int hh_level_nodes(hh_handle_t h, int level, int64_t* out) {
    if (!h || !out) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        (h->lows.empty() ? h->subs[0] : h->lows[0])->level_nodes(level, out);
        return HH_OK;
    });
}

int hh_get_level_stencil(hh_handle_t h, int level, void* coef_out) {
    if (!h || !coef_out) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(h->lows.empty(), HH_ERR_UNSUPPORTED, "hh_get_level_stencil: not available on a mixed-precision handle");
        HH_REQUIRE(h->slab_mode == 0, HH_ERR_UNSUPPORTED, "hh_get_level_stencil: use hh_slab_level_stencil on a slab handle");
        h->subs[0]->get_level_stencil(level, coef_out);
        return HH_OK;
    });
}

// Galerkin stencil of one slab (local planes, halo planes included): parity hook for MGsetup under slab decomposition
int hh_slab_level_stencil(hh_handle_t h, int slab, int level, int64_t* n_local_out, void* coef_out) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(h->lows.empty(), HH_ERR_UNSUPPORTED, "hh_slab_level_stencil: not available on a mixed-precision handle");
        HH_REQUIRE(slab >= 0 && slab < (int)h->subs.size(), HH_ERR_ARG, "hh_slab_level_stencil: bad slab index");
        if (n_local_out) h->subs[slab]->level_nodes(level, n_local_out);
        if (coef_out) h->subs[slab]->get_level_stencil(level, coef_out);
        return HH_OK;
    });
}

int hh_get_diagonal(hh_handle_t h, int shifted, double shift, double* diag_out) {
    if (!h || !diag_out) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        if (h->slab_mode == 0) {
            h->subs[0]->get_diagonal(shifted, shift, diag_out);
            return HH_OK;
        }
        // slab handle: every slab evaluates its planes; the caller's array holds the planes its B / X hold
        const int64_t plane = (int64_t)h->pb.n[0] * h->pb.n[1];
        for (auto& sb : h->subs) {
            const SlabLevel& g = sb->sgeo[0];
            std::vector<double> loc((size_t)2 * plane * g.nloc);
            sb->get_diagonal(shifted, shift, loc.data());
            const int64_t dst0 = h->slab_mode == 1 ? g.own0 : 0;
            std::memcpy(diag_out + 2 * plane * dst0, loc.data() + 2 * plane * g.zb, sizeof(double) * 2 * plane * (g.own1 - g.own0));
        }
        return HH_OK;
    });
}

// ---------------------------------------------------------------------------------------------
int hh_apply_device(hh_handle_t h, const void* dX, void* dY, int64_t nrhs, int shifted, double shift, int transpose) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(dX && dY && nrhs >= 1 && dX != dY, HH_ERR_ARG, "hh_apply_device: bad arguments");
        HH_REQUIRE(h->subs.size() == 1, HH_ERR_UNSUPPORTED, "device-pointer entry points need a single-device handle");
        h->subs[0]->apply_device(dX, dY, nrhs, shifted, shift, transpose);
        return HH_OK;
    });
}

// right-hand sides replica i can hold at once next to its work vectors (and its ComplexF32 companion's, if mixed)
static int64_t batch_limit(hh_handle_t h, int i, const hh_solve_options& o) {
    SolverBase* s = h->subs[i].get();
    // the staging slots of earlier hh_solve calls are kept and reused: they count as available
    double staged = 0.0;
    for (int q = 0; q < 2; ++q) staged += (double)h->stages[i]->db[q].n + (double)h->stages[i]->dx[q].n;
    if (h->lows.empty()) return s->max_rhs_per_batch(o, staged);
    SolverBase* lo = h->lows[i].get();
    HH_CUDA(cudaSetDevice(s->device));
    size_t fr = 0, tot = 0;
    HH_CUDA(cudaMemGetInfo(&fr, &tot));
    const double staging = 2.0 * (double)lo->internal_ld() * sizeof(cx<float>);
    const double per = s->per_rhs_bytes(o) + lo->cycle_bytes_per_rhs() + staging;
    const double held = s->held_bytes() + lo->held_bytes() + (double)(h->lo_b[i]->n + h->lo_z[i]->n) * sizeof(cx<float>) + staged;
    return std::max<int64_t>((int64_t)std::floor(0.90 * ((double)fr + held) / per), 0);
}

// ---- slab handles: how replica i sees the caller's host arrays ----
// ld: elements between consecutive right-hand sides, off: offset of the slab's first owned plane, nown: owned nodes
static void slab_host_view(hh_handle_t h, int i, int64_t& ld, int64_t& off, int64_t& nown) {
    const SolverBase* s = h->subs[i].get();
    nown = s->caller_N();
    if (h->slab_mode == 1) {
        ld = h->pb.N();
        off = (int64_t)s->sgeo[0].own0 * h->pb.n[0] * h->pb.n[1];
    } else {
        ld = nown;
        off = 0;
    }
}
// nodes per right-hand side of the device blocks handed to the *_device entry points
static int64_t device_block_N(hh_handle_t h) { return h->slab_mode == 2 ? h->subs[0]->caller_N() : h->pb.N(); }

// right-hand sides per batch that every slab can hold (the slabs solve in lockstep, so they must agree)
static int64_t slab_batch_limit(hh_handle_t h, const hh_solve_options& o) {
    int64_t k = INT64_MAX;
    for (size_t i = 0; i < h->subs.size(); ++i) k = std::min(k, batch_limit(h, (int)i, o));
    if (h->slab_mode == 2) {
        SolverBase* s = h->subs[0].get();
        DevBuf<double> d;
        d.alloc(1);
        double v = (double)k;
        HH_CUDA(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, s->stream));
        s->slab->allreduce(s->stream, d.p, 1, true);
        HH_CUDA(cudaMemcpyAsync(&v, d.p, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        HH_CUDA(cudaStreamSynchronize(s->stream));
        k = (int64_t)v;
    }
    return k;
}

// k columns of `rows` elements between a strided host block and a dense device block (one copy per column: pitches of
// whole-grid blocks exceed what cudaMemcpy2D accepts)
static void copy_columns(bool to_device, char* dev, const char* host, size_t rows_bytes, size_t host_ld_bytes, int64_t k,
                         cudaStream_t st) {
    for (int64_t r = 0; r < k; ++r) {
        char* d = dev + (size_t)r * rows_bytes;
        const char* hp = host + (size_t)r * host_ld_bytes;
        if (to_device) HH_CUDA(cudaMemcpyAsync(d, hp, rows_bytes, cudaMemcpyHostToDevice, st));
        else HH_CUDA(cudaMemcpyAsync(const_cast<char*>(hp), d, rows_bytes, cudaMemcpyDeviceToHost, st));
    }
}

// host-pointer apply / solve on a slab handle: every replica works on all right-hand sides and its own planes
static void slab_apply_host(hh_handle_t h, const void* X, void* Y, int64_t nrhs, int shifted, double shift, int transpose) {
    for_each_sub(h, [&](int i) {
        SolverBase* s = h->subs[i].get();
        HH_CUDA(cudaSetDevice(s->device));
        int64_t ld, off, nown;
        slab_host_view(h, i, ld, off, nown);
        const size_t es = s->elem_size();
        DevBuf<char> dx, dy;
        dx.alloc((size_t)nown * nrhs * es);
        dy.alloc((size_t)nown * nrhs * es);
        copy_columns(true, dx.p, (const char*)X + (size_t)off * es, (size_t)nown * es, (size_t)ld * es, nrhs, s->stream);
        s->apply_device(dx.p, dy.p, nrhs, shifted, shift, transpose);
        copy_columns(false, dy.p, (const char*)Y + (size_t)off * es, (size_t)nown * es, (size_t)ld * es, nrhs, s->stream);
        HH_CUDA(cudaStreamSynchronize(s->stream));
    });
}

static int slab_solve_host(hh_handle_t h, const void* B, const int64_t* idx, const double* val, void* X, int64_t nrhs,
                           const hh_solve_options* opts, int32_t* iters_out, double* relres_out) {
    int64_t kmax = slab_batch_limit(h, *opts);
    {   // room for the staging blocks of B and X next to the work vectors
        const double kv = (opts->krylov == HH_KRYLOV_GMRES ? 2 * opts->inner + 1 : 7) + 3.5;
        kmax = (int64_t)((double)kmax * kv / (kv + 2.0));
    }
    HH_REQUIRE(kmax >= 1, HH_ERR_ALLOC, "not enough device memory for one right-hand side");
    kmax = std::min(kmax, nrhs);
    const int nsub = (int)h->subs.size();
    std::vector<int> rcs(nsub, HH_OK);
    for_each_sub(h, [&](int i) {
        SolverBase* s = h->subs[i].get();
        HH_CUDA(cudaSetDevice(s->device));
        int64_t ld, off, nown;
        slab_host_view(h, i, ld, off, nown);
        const size_t es = s->elem_size();
        DevBuf<char> db, dx;
        db.alloc((size_t)nown * kmax * es);
        dx.alloc((size_t)nown * kmax * es);
        std::vector<int64_t> idx0;
        for (int64_t c = 0; c < nrhs; c += kmax) {
            const int64_t k = std::min(kmax, nrhs - c);
            if (idx) {
                idx0.resize(k);
                for (int64_t r = 0; r < k; ++r) idx0[r] = idx[c + r] - 1;
                s->scatter_point_sources(db.p, idx0.data(), val + 2 * c, k);
            } else {
                copy_columns(true, db.p, (const char*)B + ((size_t)c * ld + off) * es, (size_t)nown * es, (size_t)ld * es, k, s->stream);
            }
            // the slabs compute identical iteration counts / residuals: the first one reports them
            const bool rep = (i == 0);
            int r = s->solve_device(db.p, dx.p, k, *opts, (rep && iters_out) ? iters_out + c : nullptr,
                                    (rep && relres_out) ? relres_out + c : nullptr);
            rcs[i] = std::max(rcs[i], r);
            copy_columns(false, dx.p, (const char*)X + ((size_t)c * ld + off) * es, (size_t)nown * es, (size_t)ld * es, k, s->stream);
            HH_CUDA(cudaStreamSynchronize(s->stream));
        }
    });
    int rc = HH_OK;
    for (int r : rcs) rc = std::max(rc, r);
    return rc;
}

// split [0,nrhs) into contiguous column ranges, one per replica
static void column_range(int64_t nrhs, int nparts, int part, int64_t& c0, int64_t& c1) {
    const int64_t base = nrhs / nparts, rem = nrhs % nparts;
    c0 = part * base + std::min<int64_t>(part, rem);
    c1 = c0 + base + (part < rem ? 1 : 0);
}

int hh_apply(hh_handle_t h, const void* X, void* Y, int64_t nrhs, int shifted, double shift, int transpose) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(X && Y && nrhs >= 1, HH_ERR_ARG, "hh_apply: bad arguments");
        if (h->slab_mode) {
            slab_apply_host(h, X, Y, nrhs, shifted, shift, transpose);
            return HH_OK;
        }
        const int64_t N = h->pb.N();
        for_each_sub(h, [&](int i) {
            SolverBase* s = h->subs[i].get();
            int64_t c0, c1;
            column_range(nrhs, (int)h->subs.size(), i, c0, c1);
            if (c1 <= c0) return;
            HH_CUDA(cudaSetDevice(s->device));
            const size_t es = s->elem_size();
            // chunk so that two N x k buffers fit comfortably
            size_t fr = 0, tot = 0;
            HH_CUDA(cudaMemGetInfo(&fr, &tot));
            int64_t kmax = std::max<int64_t>(1, (int64_t)(0.4 * (double)fr / ((double)N * es)));
            for (int64_t c = c0; c < c1; c += kmax) {
                const int64_t k = std::min(kmax, c1 - c);
                DevBuf<char> dx, dy;
                dx.alloc((size_t)N * k * es);
                dy.alloc((size_t)N * k * es);
                HH_CUDA(cudaMemcpyAsync(dx.p, (const char*)X + (size_t)c * N * es, (size_t)N * k * es, cudaMemcpyHostToDevice, s->stream));
                s->apply_device(dx.p, dy.p, k, shifted, shift, transpose);
                HH_CUDA(cudaMemcpyAsync((char*)Y + (size_t)c * N * es, dy.p, (size_t)N * k * es, cudaMemcpyDeviceToHost, s->stream));
                HH_CUDA(cudaStreamSynchronize(s->stream));
            }
        });
        return HH_OK;
    });
}

int hh_cycle_device(hh_handle_t h, const void* dB, void* dZ, int64_t nrhs) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(dB && dZ && nrhs >= 1 && dB != dZ, HH_ERR_ARG, "hh_cycle_device: bad arguments");
        HH_REQUIRE(h->slab_mode == 0, HH_ERR_UNSUPPORTED, "hh_cycle_device is not available on a slab handle");
        HH_REQUIRE(h->subs.size() == 1, HH_ERR_UNSUPPORTED, "device-pointer entry points need a single-device handle");
        h->subs[0]->cycle_device(dB, dZ, nrhs);
        return HH_OK;
    });
}

int hh_cycle(hh_handle_t h, const void* B, void* Z, int64_t nrhs) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(B && Z && nrhs >= 1, HH_ERR_ARG, "hh_cycle: bad arguments");
        HH_REQUIRE(h->slab_mode == 0, HH_ERR_UNSUPPORTED, "hh_cycle is not available on a slab handle");
        SolverBase* s = h->subs[0].get();
        HH_CUDA(cudaSetDevice(s->device));
        const int64_t N = h->pb.N();
        const size_t es = s->elem_size();
        DevBuf<char> db, dz;
        db.alloc((size_t)N * nrhs * es);
        dz.alloc((size_t)N * nrhs * es);
        HH_CUDA(cudaMemcpy(db.p, B, (size_t)N * nrhs * es, cudaMemcpyHostToDevice));
        s->cycle_device(db.p, dz.p, nrhs);
        HH_CUDA(cudaMemcpy(Z, dz.p, (size_t)N * nrhs * es, cudaMemcpyDeviceToHost));
        return HH_OK;
    });
}

// ---------------------------------------------------------------------------------------------
static void check_solve_opts(const hh_solve_options* o) {
    HH_REQUIRE(o != nullptr, HH_ERR_ARG, "solve options are NULL");
    HH_REQUIRE(o->rel_tol >= 0.0, HH_ERR_ARG, "rel_tol must be >= 0");
}

int hh_solve_device(hh_handle_t h, const void* dB, void* dX, int64_t nrhs, const hh_solve_options* opts,
                    int32_t* iters_out, double* relres_out) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(dB && dX && nrhs >= 1 && dB != dX, HH_ERR_ARG, "hh_solve_device: bad arguments");
        check_solve_opts(opts);
        HH_REQUIRE(h->subs.size() == 1, HH_ERR_UNSUPPORTED, "device-pointer entry points need a single-device handle");
        SolverBase* s = h->subs[0].get();
        const int64_t kmax = h->slab_mode ? slab_batch_limit(h, *opts) : batch_limit(h, 0, *opts);
        HH_REQUIRE(kmax >= 1, HH_ERR_ALLOC, "not enough device memory for one right-hand side");
        const int64_t N = device_block_N(h);
        const size_t es = s->elem_size();
        int rc = HH_OK;
        for (int64_t c = 0; c < nrhs; c += kmax) {
            const int64_t k = std::min(kmax, nrhs - c);
            int r = s->solve_device((const char*)dB + (size_t)c * N * es, (char*)dX + (size_t)c * N * es, k, *opts,
                                    iters_out ? iters_out + c : nullptr, relres_out ? relres_out + c : nullptr);
            rc = std::max(rc, r);
        }
        return rc;
    });
}

// host-pointer solve; `idx`/`val` non-NULL selects point-source right-hand sides instead of B
static int solve_host(hh_handle_t h, const void* B, const int64_t* idx, const double* val, void* X, int64_t nrhs,
                      const hh_solve_options* opts, int32_t* iters_out, double* relres_out) {
    check_solve_opts(opts);
    const int64_t N = h->pb.N();
    const int nsub = (int)h->subs.size();
    std::vector<int> rcs(nsub, HH_OK);
    if (idx)
        for (int64_t r = 0; r < nrhs; ++r)
            HH_REQUIRE(idx[r] >= 1 && idx[r] <= N, HH_ERR_ARG, "point source index out of range (1-based)");
    if (h->slab_mode) return slab_solve_host(h, B, idx, val, X, nrhs, opts, iters_out, relres_out);
    for_each_sub(h, [&](int i) {
        SolverBase* s = h->subs[i].get();
        int64_t c0, c1;
        column_range(nrhs, nsub, i, c0, c1);
        if (c1 <= c0) return;
        HH_CUDA(cudaSetDevice(s->device));
        const size_t es = s->elem_size();
        HostStage& hs = *h->stages[i];
        hs.init();
        // Sub-batch size: what fits next to the Krylov / multigrid work vectors with two staging slots of (B, X);
        // a range that would fit in one batch is still split in two so that the PCIe copies of one half overlap
        // the solve of the other (HH_HOST_PIPELINE=0 disables the split).
        const int64_t ncols = c1 - c0;
        int64_t kmax = batch_limit(h, i, *opts);
        {
            const double kv = (opts->krylov == HH_KRYLOV_GMRES ? 2 * opts->inner + 1 : 7) + 3.5;
            kmax = (int64_t)((double)kmax * kv / (kv + 4.0));
        }
        kmax = std::min<int64_t>(std::max<int64_t>(kmax, 1), ncols);
        const char* pe = getenv("HH_HOST_PIPELINE");
        const bool pipeline = !(pe && pe[0] == '0');
        // Sub-batches.  Only the H2D copy of the FIRST and the D2H copy of the LAST sub-batch are exposed (the others
        // overlap a solve): a block of >= 8 columns is cut in two halves (a batch of 8 costs ~2 % more per column than
        // one of 16, a batch of 4 ~5 %, a batch of 2 ~45 %: finer cuts lose more than the shorter exposed copies gain).
        std::vector<int64_t> sizes;
        // Quarter / half / quarter instead of two halves: for a block of 16 columns measured as no gain at 1 GPU and 2-4 %
        // slower at 8 GPUs (profiles/bench_r02_n8_c128*.json: two batches of 4 cost what the shorter exposed copies save);
        // from 32 columns on the quarters are batches of >= 8 (~2 % dearer per column) and it is the default.
        // HH_HOST_CHUNKS=2 / 3 force either cut.
        const char* hc = getenv("HH_HOST_CHUNKS");
        const bool three = hc ? hc[0] == '3' : ncols >= 32;
        if (pipeline && three && ncols >= 16 && kmax >= (ncols + 1) / 2) {
            const int64_t q = std::max<int64_t>(4, (ncols / 4) / 4 * 4);
            sizes = {q, ncols - 2 * q, q};
        } else if (pipeline && ncols >= 8 && kmax >= (ncols + 1) / 2) {
            sizes = {(ncols + 1) / 2, ncols / 2};
        } else {
            for (int64_t c = 0; c < ncols; c += kmax) sizes.push_back(std::min(kmax, ncols - c));
        }
        int64_t kb = 0;
        std::vector<int64_t> start(sizes.size());
        for (size_t q = 0; q < sizes.size(); ++q) {
            start[q] = q ? start[q - 1] + sizes[q - 1] : c0;
            kb = std::max(kb, sizes[q]);
        }
        hs.ensure((size_t)N * kb * es);
        const int64_t nb = (int64_t)sizes.size();
        std::vector<int64_t> idx0;
        auto stage_in = [&](int64_t bidx) {  // enqueue the right-hand sides of sub-batch bidx into its slot
            const int slot = (int)(bidx & 1);
            const int64_t c = start[bidx], k = sizes[bidx];
            if (idx) {
                idx0.resize(k);
                for (int64_t r = 0; r < k; ++r) idx0[r] = idx[c + r] - 1;
                s->scatter_point_sources(hs.db[slot].p, idx0.data(), val + 2 * c, k);  // on the compute stream
            } else {
                HH_CUDA(cudaMemcpyAsync(hs.db[slot].p, (const char*)B + (size_t)c * N * es, (size_t)N * k * es,
                                        cudaMemcpyHostToDevice, hs.cs));
                HH_CUDA(cudaEventRecord(hs.h2d[slot], hs.cs));
            }
        };
        // HH_HOST_TRACE=1: wall-clock marks of the host pipeline on stderr (where an e2e call spends its time)
        const char* tr = getenv("HH_HOST_TRACE");
        const bool trace = tr && tr[0] == '1';
        const auto tr0 = std::chrono::steady_clock::now();
        auto mark = [&](const char* what, int64_t a, int64_t b) {
            if (!trace) return;
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count();
            fprintf(stderr, "[hh_solve dev %d] %9.2f ms  %s %lld %lld\n", s->device, ms, what, (long long)a, (long long)b);
        };
        mark("begin: columns, kmax", ncols, kmax);
        stage_in(0);
        for (int64_t bidx = 0; bidx < nb; ++bidx) {
            const int slot = (int)(bidx & 1);
            const int64_t c = start[bidx], k = sizes[bidx];
            mark("sub-batch enqueue: index, size", bidx, k);
            // prefetch the next sub-batch: its B slot was last read by solve(bidx-1), which has returned
            if (bidx + 1 < nb) stage_in(bidx + 1);
            if (!idx) HH_CUDA(cudaStreamWaitEvent(s->stream, hs.h2d[slot], 0));
            if (bidx >= 2) HH_CUDA(cudaStreamWaitEvent(s->stream, hs.d2h[slot], 0));  // X slot still being copied out?
            int r = s->solve_device(hs.db[slot].p, hs.dx[slot].p, k, *opts, iters_out ? iters_out + c : nullptr,
                                    relres_out ? relres_out + c : nullptr);
            rcs[i] = std::max(rcs[i], r);
            mark("sub-batch solved (host returned): index, rc", bidx, r);
            HH_CUDA(cudaEventRecord(hs.solved[slot], s->stream));
            HH_CUDA(cudaStreamWaitEvent(hs.cs, hs.solved[slot], 0));
            HH_CUDA(cudaMemcpyAsync((char*)X + (size_t)c * N * es, hs.dx[slot].p, (size_t)N * k * es, cudaMemcpyDeviceToHost, hs.cs));
            HH_CUDA(cudaEventRecord(hs.d2h[slot], hs.cs));
        }
        HH_CUDA(cudaStreamSynchronize(hs.cs));
        HH_CUDA(cudaStreamSynchronize(s->stream));
        mark("end: copies drained", nb, 0);
    });
    int rc = HH_OK;
    for (int r : rcs) rc = std::max(rc, r);
    return rc;
}

int hh_solve(hh_handle_t h, const void* B, void* X, int64_t nrhs, const hh_solve_options* opts, int32_t* iters_out,
             double* relres_out) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(B && X && nrhs >= 1, HH_ERR_ARG, "hh_solve: bad arguments");
        return solve_host(h, B, nullptr, nullptr, X, nrhs, opts, iters_out, relres_out);
    });
}

int hh_solve_point_sources(hh_handle_t h, const int64_t* idx, const double* val, int64_t nrhs, void* X,
                           const hh_solve_options* opts, int32_t* iters_out, double* relres_out) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        HH_REQUIRE(idx && val && X && nrhs >= 1, HH_ERR_ARG, "hh_solve_point_sources: bad arguments");
        return solve_host(h, nullptr, idx, val, X, nrhs, opts, iters_out, relres_out);
    });
}

// ---------------------------------------------------------------------------------------------
int hh_get_counters(hh_handle_t h, double* setup_seconds, double* solve_seconds, int64_t* n_prec,
                    int64_t* kernel_launches) {
    if (!h) return HH_ERR_ARG;
    double a = 0, b = 0;
    int64_t c = 0, d = 0;
    for (auto& s : h->subs) {
        a = std::max(a, s->setup_seconds);
        b = std::max(b, s->solve_seconds);
        c += s->n_prec;
        d += s->launches;
    }
    for (auto& s : h->lows) {
        a = std::max(a, s->setup_seconds);
        d += s->launches;
    }
    if (setup_seconds) *setup_seconds = a;
    if (solve_seconds) *solve_seconds = b;
    if (n_prec) *n_prec = c;
    if (kernel_launches) *kernel_launches = d;
    return HH_OK;
}

int hh_profile_enable(hh_handle_t h, int on) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        for (auto* v : {&h->subs, &h->lows})
            for (auto& s : *v) {
                HH_CUDA(cudaSetDevice(s->device));
                if (!on) s->prof.flush(s->stream);
                s->prof.on = on != 0;
            }
        return HH_OK;
    });
}

int hh_profile_reset(hh_handle_t h) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        for (auto* v : {&h->subs, &h->lows})
            for (auto& s : *v) {
                HH_CUDA(cudaSetDevice(s->device));
                s->prof.flush(s->stream);
                s->prof.reset();
            }
        return HH_OK;
    });
}

int hh_profile_num_tags(void) { return T_NTAGS; }

const char* hh_profile_tag_name(int tag) { return (tag >= 0 && tag < T_NTAGS) ? kTagNames[tag] : ""; }

int hh_profile_get(hh_handle_t h, int tag, int64_t* launches, double* milliseconds, double* algorithmic_bytes) {
    if (!h || tag < 0 || tag >= T_NTAGS) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        int64_t n = 0;
        double ms = 0, by = 0;
        for (auto* v : {&h->subs, &h->lows})
            for (auto& s : *v) {
                HH_CUDA(cudaSetDevice(s->device));
                s->prof.flush(s->stream);
                n += s->prof.count[tag];
                ms += s->prof.ms[tag];
                by += s->prof.bytes[tag];
            }
        if (launches) *launches = n;
        if (milliseconds) *milliseconds = ms;
        if (algorithmic_bytes) *algorithmic_bytes = by;
        return HH_OK;
    });
}

// entries = launches with identical work (same kernel class and algorithmic bytes per launch)
int hh_profile_num_entries(hh_handle_t h) {
    if (!h) return 0;
    try {
        SolverBase* s = h->subs[0].get();
        cudaSetDevice(s->device);
        s->prof.flush(s->stream);
        int n = (int)s->prof.entries.size();
        if (!h->lows.empty()) {
            h->lows[0]->prof.flush(h->lows[0]->stream);
            n += (int)h->lows[0]->prof.entries.size();
        }
        return n;
    } catch (...) {
        return 0;
    }
}

int hh_profile_entry(hh_handle_t h, int index, int* tag, int64_t* launches, double* milliseconds,
                     double* algorithmic_bytes_per_launch) {
    if (!h) return HH_ERR_ARG;
    return guarded(h, [&]() -> int {
        SolverBase* s = h->subs[0].get();
        const int n0 = (int)s->prof.entries.size();
        const int n1 = h->lows.empty() ? 0 : (int)h->lows[0]->prof.entries.size();
        HH_REQUIRE(index >= 0 && index < n0 + n1, HH_ERR_ARG, "hh_profile_entry: bad index");
        const auto& e = index < n0 ? s->prof.entries[index] : h->lows[0]->prof.entries[index - n0];
        if (tag) *tag = e.tag;
        if (launches) *launches = e.count;
        if (milliseconds) *milliseconds = e.ms;
        if (algorithmic_bytes_per_launch) *algorithmic_bytes_per_launch = e.bytes / (double)e.count;
        return HH_OK;
    });
}

}  // extern "C"
