"""CPU: the tiling of the implicit operator's TMA kernel (k_spmv_dot_tmac) as the HOST builds it -- tile descriptors
with two lane shapes, 7 or 8 nodes per thread, the pure row block + keep mask of every chunk, the list of nodes that
k_spmv_fix serves -- replayed with the kernels' own ownership rules: every interior node must be produced exactly once,
with the row block the node really has; every shared-memory index must stay inside the brick; the Ap box of a tile
(TMA store) must be a 16-byte multiple wide, fit in the brick's memory and never reach another tile's nodes.  (The GPU tests check bit-identity of
the operator on about 20 shapes; this covers the bench sizes, 200^3 and odd shapes without a GPU.)"""
import ctypes as C

import numpy as np
import pytest

import micropp_b200 as M

TILE_Y, TILE_Z = 8, 4


def sphere_types(nx, ny, nz, r=0.2):
    c = [(np.arange(n - 1) + 0.5) / (n - 1) for n in (nx, ny, nz)]
    X, Y, Z = np.meshgrid(*c, indexing="ij")
    t = (((X - .5) ** 2 + (Y - .5) ** 2 + (Z - .5) ** 2) < r * r).astype(np.int32)     # [ex, ey, ez]
    return np.ascontiguousarray(t.transpose(2, 1, 0)).reshape(-1)                          # e = (ez*ney + ey)*nex + ex


def layer_types(nx, ny, nz):
    t = np.zeros((nz - 1, ny - 1, nx - 1), dtype=np.int32)
    t[:, (ny - 1) // 2:, :] = 1
    t[: (nz - 1) // 3, :, :] = 2
    return t.reshape(-1)


def tiling(lib, dims, et):
    nx, ny, nz = dims
    ip = C.POINTER(C.c_int)
    f = lib.mgpu_tmac_tiling_host
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 3 + [ip] * 6
    meta = np.zeros(6, dtype=np.int32)
    et = np.ascontiguousarray(et, dtype=np.int32)
    nrows = f(nx, ny, nz, et.ctypes.data_as(ip), meta.ctypes.data_as(ip), None, None, None, None)
    tn, cb, nchunk, pitch, ntiles, nfix = (int(v) for v in meta)
    nix, niy, niz = nx - 2, ny - 2, nz - 2
    rowid = np.zeros(nix * niy * niz, dtype=np.int32)
    tiles = np.zeros((ntiles, 4), dtype=np.int32)
    pure = np.zeros(niz * niy * nchunk, dtype=np.int32)
    fix = np.zeros((max(nfix, 1), 2), dtype=np.int32)
    f(nx, ny, nz, et.ctypes.data_as(ip), meta.ctypes.data_as(ip), rowid.ctypes.data_as(ip), tiles.ctypes.data_as(ip),
      pure.ctypes.data_as(ip), fix.ctypes.data_as(ip))
    return dict(tn=tn, cb=cb, nchunk=nchunk, pitch=pitch, ntiles=ntiles, nfix=nfix, nrows=nrows,
                rowid=rowid.reshape(niz, niy, nix), tiles=tiles, pure=pure.reshape(niz, niy, nchunk), fix=fix[:nfix])


@pytest.mark.parametrize("dims,kind", [((30, 30, 30), "sphere"), ((50, 50, 50), "sphere"), ((40, 40, 40), "layer"),
                                        ((40, 40, 40), "sphere"), ((200, 200, 28), "sphere"), ((3, 3, 3), "sphere"),
                                        ((19, 12, 5), "layer"), ((66, 5, 6), "sphere"), ((11, 10, 13), "layer"),
                                        ((24, 15, 9), "sphere"), ((16, 14, 21), "layer"), ((10, 6, 14), "sphere"),
                                        ((34, 7, 8), "layer"), ((48, 11, 6), "sphere")])
def test_tmac_tiling_covers_every_interior_node_once(dims, kind):
    lib = M.load()
    nx, ny, nz = dims
    et = sphere_types(*dims) if kind == "sphere" else layer_types(*dims)
    t = tiling(lib, dims, et)
    nix, niy, niz = nx - 2, ny - 2, nz - 2
    tn, cb, pitch = t["tn"], t["cb"], t["pitch"]
    assert tn in (7, 8) and 1 <= cb <= 4 and t["nchunk"] == -(-nix // tn)
    pxo = tn * cb - 2                                # the Ap box of a tile: its middle nodes (16-B aligned TMA origin)
    assert pxo % 2 == 0 and pxo >= 2 and 3 * 32 * pxo <= 3 * 60 * pitch
    assert pitch % 2 == 0 and (pitch // 2) % 2 == 1 and pitch >= tn * cb + 2      # 16-B rows, conflict-free quarter-warps
    assert t["nrows"] >= 3 and t["rowid"].max() < t["nrows"]
    count = np.zeros((niz, niy, nix), dtype=np.int32)
    ii = np.arange(nix)
    for tile in range(t["ntiles"]):
        c0, y0, z0, shape = (int(v) for v in t["tiles"][tile])
        by, bz = (TILE_Z + 2, TILE_Y + 2) if shape else (TILE_Y + 2, TILE_Z + 2)       # brick rows in y / z
        ny_t, nz_t = (TILE_Z, TILE_Y) if shape else (TILE_Y, TILE_Z)
        xoff = (c0 * tn) & 1                                                           # TMA box starts at an even x
        assert ((c0 * tn) & ~1) % 2 == 0
        for w in range(cb):
            xs = tn * w + xoff
            assert (xs & ~1) + 10 <= pitch                                             # five 16-B loads stay inside the row
            c = c0 + w
            if c >= t["nchunk"]:
                continue
            nvalid = min(tn, nix - c * tn)
            jj = np.arange(y0, min(y0 + ny_t, niy))
            kk = np.arange(z0, min(z0 + nz_t, niz))
            info = t["pure"][np.ix_(kk, jj)][:, :, c]
            keep = ((1 << nvalid) - 1) & ~(info >> 8)
            for tt in range(nvalid):
                sel = ((keep >> tt) & 1).astype(bool)
                node_ids = t["rowid"][np.ix_(kk, jj)][:, :, c * tn + tt]
                assert np.all(node_ids[sel] == (info & 0xff)[sel])                     # kept nodes use the chunk's pure block
                count[np.ix_(kk, jj, [c * tn + tt])][:, :, 0]                          # (bounds check of the index)
                count[kk[:, None], jj[None, :], c * tn + tt] += sel
    # tiles partition the chunk grid: no chunk is served twice, so the Ap boxes (which may only overhang the interior at
    # its high ends, where they meet boundary nodes or leave the grid) never touch another tile's nodes
    owner = np.zeros((niz, niy, t["nchunk"]), dtype=np.int32)
    for tile in range(t["ntiles"]):
        c0, y0, z0, shape = (int(v) for v in t["tiles"][tile])
        ny_t, nz_t = (TILE_Z, TILE_Y) if shape else (TILE_Y, TILE_Z)
        owner[z0:z0 + nz_t, y0:y0 + ny_t, c0:c0 + cb] += 1
    assert owner.min() == 1 and owner.max() == 1
    # the nodes no chunk keeps go to k_spmv_fix with their own row block
    for node, rid in t["fix"]:
        k, rem = divmod(int(node), nx * ny)
        j, i = divmod(rem, nx)
        assert 1 <= i <= nix and 1 <= j <= niy and 1 <= k <= niz
        assert rid == t["rowid"][k - 1, j - 1, i - 1]
        count[k - 1, j - 1, i - 1] += 1
    assert count.min() == 1 and count.max() == 1
    assert t["nfix"] >= int(np.sum(t["rowid"] >= 3))       # interface nodes + minority nodes of a chunk
    if kind == "sphere" and min(dims) >= 24:
        assert t["nfix"] < 0.08 * nix * niy * niz


def test_tmac_tiling_efficiency_at_bench_sizes():
    """Executed node slots per interior node: 30^3 -> 22400 / 21952 (7 nodes per thread + a 4y x 8z strip), 50^3 exact;
    40^3 pays for the even-width Ap box (7 nodes per thread only with 2 or 4 warps): 8 nodes x 3 warps, 48 slots for 38."""
    lib = M.load()
    for n, bound in ((30, 1.03), (50, 1.0001), (40, 1.42), (200, 1.12)):
        dims = (n, n, 12 if n == 200 else n)
        t = tiling(lib, dims, sphere_types(*dims))
        executed = t["ntiles"] * t["cb"] * 32 * t["tn"]
        assert executed <= bound * (dims[0] - 2) * (dims[1] - 2) * (dims[2] - 2) * (1.35 if n == 200 else 1.0), (n, executed)
