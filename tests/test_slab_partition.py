"""Slab decomposition of one grid over several GPUs: the host-side partition
(hh_slab_partition, include/helmholtz_b200.h).  CPU only: no compute entry point is called."""
import itertools

import pytest


@pytest.mark.parametrize("n3,levels,nranks", [(513, 3, 8), (257, 3, 8), (257, 3, 2), (129, 3, 4), (65, 3, 3), (65, 4, 2),
                                              (33, 2, 5), (33, 3, 8), (17, 1, 4), (129, 5, 8)])
def test_partition_covers_aligns_and_nests(pkg, n3, levels, nranks):
    parts = [pkg.slabPartition(n3, levels, nranks, r) for r in range(nranks)]
    for l in range(levels):
        f = 1 << (levels - 1 - l)
        n2g = (n3 - 1) // (1 << l) + 1
        # owned ranges tile the level's planes in rank order
        assert parts[0][l]["own0"] == 0 and parts[-1][l]["own1"] == n2g
        for a, b in zip(parts[:-1], parts[1:]):
            assert a[l]["own1"] == b[l]["own0"]
        for r, p in enumerate(parts):
            g = p[l]
            assert g["n2g"] == n2g
            assert g["own1"] > g["own0"]
            # local numbering: plane z is global plane z + koff; owned planes are zb..ze-1
            assert g["koff"] + g["zb"] == g["own0"] and g["ze"] - g["zb"] == g["own1"] - g["own0"]
            # one halo plane below (inside the alignment planes) and above, except at the ends of the grid
            assert (g["zb"] >= 1) == (r > 0) and (g["nloc"] == g["ze"] + 1) == (r < nranks - 1)
            assert g["koff"] >= 0 and g["koff"] + g["nloc"] <= n2g
            # local plane 0 sits on an even global plane of every coarser level: z >> 1 is the global coarsening
            assert g["koff"] % f == 0
            if l + 1 < levels:
                c = p[l + 1]
                assert g["koff"] == 2 * c["koff"] and g["own0"] == 2 * c["own0"]
                # restriction of every owned coarse plane reads fine planes 2K-1..2K+1 that are held locally
                assert 2 * c["zb"] - 1 >= (0 if r else -1) and 2 * (c["ze"] - 1) + 1 <= g["nloc"] - 1 + (r == nranks - 1)
                # interpolation to every owned fine plane reads coarse planes z>>1 and (z>>1)+1 that are held locally
                assert ((g["ze"] - 1) >> 1) + ((g["ze"] - 1) & 1) <= c["nloc"] - 1
                assert (g["zb"] >> 1) >= 0
    # balanced on the coarsest level: cell counts differ by at most one
    cells = [p[-1]["own1"] - p[-1]["own0"] for p in parts]
    cells[-1] -= 1  # the last slab also owns the closing plane
    assert max(cells) - min(cells) <= 1


def test_partition_rejects_what_cannot_be_split(pkg):
    with pytest.raises(ValueError):
        pkg.slabPartition(33, 3, 9, 0)   # coarsest level has 8 cells
    with pytest.raises(ValueError):
        pkg.slabPartition(34, 2, 2, 0)   # 33 cells are not divisible by 2
    with pytest.raises(ValueError):
        pkg.slabPartition(33, 3, 2, 2)   # rank out of range


def test_single_slab_is_the_whole_grid(pkg):
    for n3, levels in itertools.product((17, 33, 65), (1, 2, 3)):
        p = pkg.slabPartition(n3, levels, 1, 0)
        for l, g in enumerate(p):
            n = (n3 - 1) // (1 << l) + 1
            assert (g["own0"], g["own1"], g["koff"], g["nloc"], g["zb"], g["ze"], g["n2g"]) == (0, n, 0, n, 0, n, n)


def test_partition_properties_randomised(pkg):
    """the same invariants over many random (cells, levels, slabs): tiling, alignment, nesting, halo sufficiency"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 5), st.integers(1, 40), st.integers(1, 12))
    def check(levels, coarse_cells, nranks):
        n3 = coarse_cells * (1 << (levels - 1)) + 1
        if coarse_cells < nranks:
            with pytest.raises(ValueError):
                pkg.slabPartition(n3, levels, nranks, 0)
            return
        parts = [pkg.slabPartition(n3, levels, nranks, r) for r in range(nranks)]
        for l in range(levels):
            n2g = (n3 - 1) // (1 << l) + 1
            assert parts[0][l]["own0"] == 0 and parts[-1][l]["own1"] == n2g
            for r, p in enumerate(parts):
                g = p[l]
                assert g["own1"] > g["own0"] and g["n2g"] == n2g
                assert g["koff"] + g["zb"] == g["own0"] and g["ze"] - g["zb"] == g["own1"] - g["own0"]
                assert 0 <= g["koff"] and g["koff"] + g["nloc"] <= n2g
                if r + 1 < nranks:
                    assert g["own1"] == parts[r + 1][l]["own0"] and g["nloc"] == g["ze"] + 1
                if r > 0:
                    assert g["zb"] >= 1
                if l + 1 < levels:
                    c = p[l + 1]
                    assert g["koff"] == 2 * c["koff"] and g["own0"] == 2 * c["own0"]
                    # fine planes read by the restriction to / written by the interpolation from the owned coarse planes
                    lo_f = 2 * c["own0"] - (1 if c["own0"] > 0 else 0)
                    hi_f = 2 * (c["own1"] - 1) + (1 if c["own1"] < c["n2g"] else 0)
                    assert g["koff"] <= lo_f and hi_f <= g["koff"] + g["nloc"] - 1
                    # coarse planes read by the interpolation to the owned fine planes
                    hi_c = (g["own1"] - 1 + 1) // 2
                    assert c["koff"] <= g["own0"] // 2 and hi_c <= c["koff"] + c["nloc"] - 1

    check()
