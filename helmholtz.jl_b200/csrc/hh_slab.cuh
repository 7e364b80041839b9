// hh_slab.cuh -- slab domain decomposition of ONE problem over several GPUs (the grid that
// exceeds one GPU, config 5): partition of the planes of the last dimension, and the two collectives the solve
// needs -- neighbour halo exchange of planes and an all-reduce of the per-RHS dot/norm partials.
//
// Two transports implement them:
//   * NcclTransport   one process per GPU (torchrun / Julia Distributed workers); ncclSend/ncclRecv groups over NVLink
//                     for the halos, ncclAllReduce for the scalars.  libnccl is bound at run time (dlopen) so that the
//                     library has no link-time dependency: inside a process that already loaded torch's NCCL that
//                     copy is used.
//   * ThreadTransport one process, one host thread per slab (the slabs may live on different devices -- peer copies
//                     over NVLink -- or share a device, which is how the single-GPU tests cover this path).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "hh_common.cuh"

namespace hh {

// (included by hh_solver.cuh after hh::Error / HH_CUDA / HH_REQUIRE are defined)

// geometry of one multigrid level inside one slab (planes of the last dimension)
struct SlabLevel {
    int own0 = 0, own1 = 0;  // global planes owned: own0 <= k < own1
    int koff = 0;            // global index of local plane 0
    int nloc = 0;            // local planes held: alignment/halo planes below + owned + one halo plane above
    int zb = 0, ze = 0;      // owned planes in local numbering
    int n2g = 0;             // global plane count of the level
};

// Partition of an n3-plane grid with `levels` multigrid levels over `nranks` slabs.  The coarsest level's cells are
// split as evenly as possible; finer levels follow by doubling (a slab that owns coarse planes [K0,K1) owns fine planes
// [2K0,2K1)), so that restriction and interpolation between a slab's own levels need a one-plane halo only.  Below the owned range a slab keeps
// 2^(levels-1-l) planes on level l (only the top one is exchanged): that keeps local plane 0 of every level at an even
// global index, i.e. local coarsening "z >> 1" agrees with the global one.  Returns false when a slab would be empty.
inline bool slab_partition(int n3, int levels, int nranks, int rank, std::vector<SlabLevel>& out) {
    out.assign(levels, SlabLevel());
    const int f0 = 1 << (levels - 1);
    if (nranks < 1 || rank < 0 || rank >= nranks || n3 < 2 || ((n3 - 1) % f0) != 0) return false;
    const int cells = (n3 - 1) / f0;  // cells of the coarsest level
    if (cells < nranks) return false;
    const int base = cells / nranks, rem = cells % nranks;
    const int c0 = rank * base + std::min(rank, rem);
    const int c1 = c0 + base + (rank < rem ? 1 : 0);
    const bool first = rank == 0, last = rank == nranks - 1;
    for (int l = 0; l < levels; ++l) {
        const int f = 1 << (levels - 1 - l);
        SlabLevel& s = out[l];
        s.n2g = cells * f + 1;
        s.own0 = c0 * f;
        s.own1 = last ? s.n2g : c1 * f;
        const int pad_lo = first ? 0 : f, pad_hi = last ? 0 : 1;
        s.koff = s.own0 - pad_lo;
        s.zb = pad_lo;
        s.ze = pad_lo + (s.own1 - s.own0);
        s.nloc = s.ze + pad_hi;
    }
    return true;
}

struct SlabTransport {
    int rank = 0, nranks = 1;
    virtual ~SlabTransport() {}
    // For each of `nseg` segments (segment r starts at base + r*stride bytes): send `bytes` bytes at offset send_dn to
    // the slab below and at send_up to the slab above; receive the neighbours' counterparts at recv_dn / recv_up.
    // Enqueued on `st` of `device`; the first / last slab skips the missing side.  fill_lower / fill_upper select which
    // halo is wanted: the lower one (data moves up: send_up -> the upper neighbour's recv_dn), the upper one, or both.
    virtual void exchange(cudaStream_t st, int device, char* base, size_t stride, int nseg, size_t bytes, size_t send_dn,
                          size_t recv_dn, size_t send_up, size_t recv_up, bool fill_lower, bool fill_upper) = 0;
    // in-place reduction of n doubles on the device over all slabs; every slab receives bit-identical results
    virtual void allreduce(cudaStream_t st, double* dbuf, int n, bool minimum) = 0;
    // true when exchange() only enqueues work on `st` (no host synchronisation): it may then overlap other streams
    virtual bool stream_ordered() const { return false; }
};

// ------------------------------------------------------------------------------------------- NCCL
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    bool ok = false;
    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] { api.load(); });
        return api;
    }
    void load() {
        const char* env = getenv("HH_NCCL_LIB");
        void* h = nullptr;
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm) continue;
            h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            error = "libnccl.so.2 could not be loaded (set HH_NCCL_LIB)";
            return;
        }
#define HH_NCCL_SYM(name)                                                          \
    name = (decltype(name))dlsym(h, "nccl" #name);                                 \
    if (!name) {                                                                   \
        error = "symbol nccl" #name " is missing from the NCCL library";           \
        return;                                                                    \
    }
        HH_NCCL_SYM(GetUniqueId)
        HH_NCCL_SYM(CommInitRank)
        HH_NCCL_SYM(CommDestroy)
        HH_NCCL_SYM(Send)
        HH_NCCL_SYM(Recv)
        HH_NCCL_SYM(AllReduce)
        HH_NCCL_SYM(GroupStart)
        HH_NCCL_SYM(GroupEnd)
        HH_NCCL_SYM(GetErrorString)
#undef HH_NCCL_SYM
        ok = true;
    }
};

// ------------------------------------------------------------------------------------------- threads
// shared state of the slabs of one in-process handle
struct ThreadGroup {
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    bool failed = false;
    struct Slot {
        char* base = nullptr;
        size_t stride = 0, send_dn = 0, send_up = 0;
        int device = 0;
        std::vector<double> red;
    };
    std::vector<Slot> slots;
    explicit ThreadGroup(int n_) : n(n_), slots(n_) {}
    // all slabs arrive, or any slab reported a failure (then every waiter throws instead of hanging)
    bool barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (failed) return false;
        const uint64_t gen = generation;
        if (++waiting == n) {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return generation != gen || failed; });
        return !failed;
    }
    void fail() {
        std::lock_guard<std::mutex> lk(mu);
        failed = true;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(mu);
        failed = false;
        waiting = 0;
    }
};


#define HH_NCCL(call)                                                                                          \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != ncclSuccess)                                                                                \
            throw hh::Error(HH_ERR_CUDA, std::string("NCCL: ") + #call + ": " + NcclApi::get().GetErrorString(r__)); \
    } while (0)

// Gather the outgoing planes of all segments into the two contiguous send blocks (pack) / scatter the two received blocks
// into the halo planes (unpack): ONE launch each instead of two strided cudaMemcpy2DAsync per direction -- the exchange is
// latency-bound (thousands of exchanges of 2-34 MB per batch), so every launch around the ncclSend/ncclRecv pair counts.
// Segment r starts at base + r*stride; block layout: segment-major, `bytes` (a multiple of 16) per segment.
__global__ void __launch_bounds__(256) k_halo_copy(char* __restrict__ base, size_t stride, int nseg, size_t bytes, char* __restrict__ blk_a,
                                                   size_t off_a, char* __restrict__ blk_b, size_t off_b, int to_block) {
    // blk_a <-> base + off_a (if blk_a), blk_b <-> base + off_b (if blk_b); to_block: planes -> blocks, else blocks -> planes
    const size_t n16 = bytes / 16;
    const size_t tot = n16 * nseg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const size_t r = t / n16, e = t - r * n16;
        if (blk_a) {
            uint4* pl = reinterpret_cast<uint4*>(base + r * stride + off_a) + e;
            uint4* bk = reinterpret_cast<uint4*>(blk_a + r * bytes) + e;
            if (to_block) *bk = *pl;
            else *pl = *bk;
        }
        if (blk_b) {
            uint4* pl = reinterpret_cast<uint4*>(base + r * stride + off_b) + e;
            uint4* bk = reinterpret_cast<uint4*>(blk_b + r * bytes) + e;
            if (to_block) *bk = *pl;
            else *pl = *bk;
        }
    }
}

struct NcclTransport : SlabTransport {
    ncclComm_t comm = nullptr;
    NcclTransport(int rank_, int nranks_, const void* unique_id) {
        rank = rank_;
        nranks = nranks_;
        NcclApi& api = NcclApi::get();
        HH_REQUIRE(api.ok, HH_ERR_UNSUPPORTED, api.error);
        ncclUniqueId id;
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        memcpy(&id, unique_id, sizeof(id));
        HH_NCCL(api.CommInitRank(&comm, nranks, id, rank));
        const char* e = getenv("HH_HALO_PACK");
        pack = !(e && e[0] == '0');
    }
    ~NcclTransport() override {
        if (comm) NcclApi::get().CommDestroy(comm);
        if (stage) cudaFree(stage);
    }
    // packed = one message per neighbour and direction: the nseg planes are gathered into / scattered from contiguous
    // staging buffers by strided device copies (HH_HALO_PACK=0 sends every plane as its own message instead)
    char* stage = nullptr;
    size_t stage_bytes = 0;
    bool pack = true;
    void exchange(cudaStream_t st, int, char* base, size_t stride, int nseg, size_t bytes, size_t send_dn, size_t recv_dn,
                  size_t send_up, size_t recv_up, bool fill_lower, bool fill_upper) override {
        NcclApi& api = NcclApi::get();
        if (nranks == 1) return;
        const bool lo = rank > 0, hi = rank < nranks - 1;
        // what this slab does: lower halo wanted -> send my top plane up, receive my lower halo from below; etc.
        const bool s_up = fill_lower && hi, r_dn = fill_lower && lo, s_dn = fill_upper && lo, r_up = fill_upper && hi;
        if (pack && nseg > 1) {
            const size_t blk = (size_t)nseg * bytes;
            if (stage_bytes < 4 * blk) {
                HH_CUDA(cudaStreamSynchronize(st));
                if (stage) cudaFree(stage);
                stage = nullptr;
                stage_bytes = 0;
                cudaError_t e = cudaMalloc((void**)&stage, 4 * blk);
                if (e != cudaSuccess) throw hh::Error(HH_ERR_ALLOC, "halo staging buffer: cudaMalloc failed");
                stage_bytes = 4 * blk;
            }
            char *o_up = stage, *o_dn = stage + blk, *i_dn = stage + 2 * blk, *i_up = stage + 3 * blk;
            const bool fast = (bytes % 16 == 0) && (stride % 16 == 0) && ((uintptr_t)base % 16 == 0) && (send_up % 16 == 0) &&
                              (send_dn % 16 == 0) && (recv_up % 16 == 0) && (recv_dn % 16 == 0);
            const unsigned nb = (unsigned)std::min<size_t>(296, (blk / 16 + 255) / 256);
            if (fast) {
                if (s_up || s_dn)
                    k_halo_copy<<<nb, 256, 0, st>>>(base, stride, nseg, bytes, s_up ? o_up : nullptr, send_up, s_dn ? o_dn : nullptr, send_dn, 1);
            } else {
                if (s_up) HH_CUDA(cudaMemcpy2DAsync(o_up, bytes, base + send_up, stride, bytes, nseg, cudaMemcpyDeviceToDevice, st));
                if (s_dn) HH_CUDA(cudaMemcpy2DAsync(o_dn, bytes, base + send_dn, stride, bytes, nseg, cudaMemcpyDeviceToDevice, st));
            }
            HH_NCCL(api.GroupStart());
            if (s_up) HH_NCCL(api.Send(o_up, blk, ncclChar, rank + 1, comm, st));
            if (r_dn) HH_NCCL(api.Recv(i_dn, blk, ncclChar, rank - 1, comm, st));
            if (s_dn) HH_NCCL(api.Send(o_dn, blk, ncclChar, rank - 1, comm, st));
            if (r_up) HH_NCCL(api.Recv(i_up, blk, ncclChar, rank + 1, comm, st));
            HH_NCCL(api.GroupEnd());
            if (fast) {
                if (r_dn || r_up)
                    k_halo_copy<<<nb, 256, 0, st>>>(base, stride, nseg, bytes, r_dn ? i_dn : nullptr, recv_dn, r_up ? i_up : nullptr, recv_up, 0);
            } else {
                if (r_dn) HH_CUDA(cudaMemcpy2DAsync(base + recv_dn, stride, i_dn, bytes, bytes, nseg, cudaMemcpyDeviceToDevice, st));
                if (r_up) HH_CUDA(cudaMemcpy2DAsync(base + recv_up, stride, i_up, bytes, bytes, nseg, cudaMemcpyDeviceToDevice, st));
            }
            return;
        }
        HH_NCCL(api.GroupStart());
        for (int r = 0; r < nseg; ++r) {
            char* seg = base + (size_t)r * stride;
            if (s_up) HH_NCCL(api.Send(seg + send_up, bytes, ncclChar, rank + 1, comm, st));
            if (r_dn) HH_NCCL(api.Recv(seg + recv_dn, bytes, ncclChar, rank - 1, comm, st));
            if (s_dn) HH_NCCL(api.Send(seg + send_dn, bytes, ncclChar, rank - 1, comm, st));
            if (r_up) HH_NCCL(api.Recv(seg + recv_up, bytes, ncclChar, rank + 1, comm, st));
        }
        HH_NCCL(api.GroupEnd());
    }
    bool stream_ordered() const override { return true; }
    void allreduce(cudaStream_t st, double* dbuf, int n, bool minimum) override {
        if (nranks == 1) return;
        HH_NCCL(NcclApi::get().AllReduce(dbuf, dbuf, (size_t)n, ncclDouble, minimum ? ncclMin : ncclSum, comm, st));
    }
};

struct ThreadTransport : SlabTransport {
    std::shared_ptr<ThreadGroup> grp;
    std::vector<double> host;
    ThreadTransport(int rank_, std::shared_ptr<ThreadGroup> g) : grp(std::move(g)) {
        rank = rank_;
        nranks = grp->n;
    }
    void sync() {
        if (!grp->barrier()) throw hh::Error(HH_ERR_STATE, "another slab of this handle failed");
    }
    // pull model: every slab publishes where its outgoing planes are, then copies its neighbours' planes into its halos
    void exchange(cudaStream_t st, int device, char* base, size_t stride, int nseg, size_t bytes, size_t send_dn,
                  size_t recv_dn, size_t send_up, size_t recv_up, bool fill_lower, bool fill_upper) override {
        if (nranks == 1) return;
        ThreadGroup::Slot& me = grp->slots[rank];
        me.base = base;
        me.stride = stride;
        me.send_dn = send_dn;
        me.send_up = send_up;
        me.device = device;
        HH_CUDA(cudaStreamSynchronize(st));  // my outgoing planes are final
        sync();
        for (int r = 0; r < nseg; ++r) {
            if (rank > 0 && fill_lower) {
                const ThreadGroup::Slot& nb = grp->slots[rank - 1];
                HH_CUDA(cudaMemcpyPeerAsync(base + (size_t)r * stride + recv_dn, device, nb.base + (size_t)r * nb.stride + nb.send_up,
                                            nb.device, bytes, st));
            }
            if (rank < nranks - 1 && fill_upper) {
                const ThreadGroup::Slot& nb = grp->slots[rank + 1];
                HH_CUDA(cudaMemcpyPeerAsync(base + (size_t)r * stride + recv_up, device, nb.base + (size_t)r * nb.stride + nb.send_dn,
                                            nb.device, bytes, st));
            }
        }
        HH_CUDA(cudaStreamSynchronize(st));
        sync();  // nobody overwrites planes a neighbour is still reading
    }
    void allreduce(cudaStream_t st, double* dbuf, int n, bool minimum) override {
        if (nranks == 1) return;
        std::vector<double>& mine = grp->slots[rank].red;
        mine.resize(n);
        HH_CUDA(cudaMemcpyAsync(mine.data(), dbuf, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        HH_CUDA(cudaStreamSynchronize(st));
        sync();
        host.assign(n, 0.0);
        for (int i = 0; i < n; ++i) {  // same order on every slab: identical results
            double a = grp->slots[0].red[i];
            for (int q = 1; q < nranks; ++q) {
                const double v = grp->slots[q].red[i];
                a = minimum ? std::min(a, v) : a + v;
            }
            host[i] = a;
        }
        sync();  // everybody has read the slots
        HH_CUDA(cudaMemcpyAsync(dbuf, host.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        HH_CUDA(cudaStreamSynchronize(st));
    }
};

}  // namespace hh
