"""CPU: the tiling of the implicit operator's TMA kernel (k_spmv_dot_tmac) as the HOST builds it -- tile descriptors
with two lane shapes, 7 or 8 nodes per thread, the pure row block + fix-up mask of every chunk, the fix-up tasks --
replayed with the kernel's own ownership rules: every interior node must be produced exactly once, with the row block
the node really has, and every shared-memory index must stay inside the brick.  (The GPU tests check bit-identity of
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
    f.argtypes = [C.c_int] * 3 + [ip] * 7
    meta = np.zeros(6, dtype=np.int32)
    et = np.ascontiguousarray(et, dtype=np.int32)
    nrows = f(nx, ny, nz, et.ctypes.data_as(ip), meta.ctypes.data_as(ip), None, None, None, None, None)
    tn, cb, nchunk, pitch, ntiles, ntasks = (int(v) for v in meta)
    nix, niy, niz = nx - 2, ny - 2, nz - 2
    rowid = np.zeros(nix * niy * niz, dtype=np.int32)
    tiles = np.zeros((ntiles, 4), dtype=np.int32)
    pure = np.zeros(niz * niy * nchunk, dtype=np.int32)
    fptr = np.zeros(ntiles + 1, dtype=np.int32)
    tasks = np.zeros((max(ntasks, 1), 4), dtype=np.int32)
    f(nx, ny, nz, et.ctypes.data_as(ip), meta.ctypes.data_as(ip), rowid.ctypes.data_as(ip), tiles.ctypes.data_as(ip),
      pure.ctypes.data_as(ip), fptr.ctypes.data_as(ip), tasks.ctypes.data_as(ip))
    return dict(tn=tn, cb=cb, nchunk=nchunk, pitch=pitch, ntiles=ntiles, ntasks=ntasks, nrows=nrows,
                rowid=rowid.reshape(niz, niy, nix), tiles=tiles, pure=pure.reshape(niz, niy, nchunk), fptr=fptr,
                tasks=tasks[:ntasks])


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
        for e in t["tasks"][t["fptr"][tile]:t["fptr"][tile + 1]]:
            for pos in (int(e[0]), int(e[1])):
                if pos < 0:
                    continue
                lx, fy, fz = pos & 0xff, (pos >> 8) & 0xf, (pos >> 12) & 0xf
                assert fy < ny_t and fz < nz_t and lx < tn * cb
                assert lx + xoff + 2 < pitch and fy + 2 < by and fz + 2 < bz           # neighbours inside the brick
                x, y, z = c0 * tn + lx, y0 + fy, z0 + fz
                assert t["rowid"][z, y, x] == int(e[2])                                # the task's row block is the node's
                count[z, y, x] += 1
    assert count.min() == 1 and count.max() == 1
    # interface nodes = nodes whose 8 elements are not all of one material
    assert t["fptr"][-1] == t["ntasks"]


def test_tmac_tiling_efficiency_at_bench_sizes():
    """Executed node slots per interior node: 30^3 -> 22400 / 21952 (7 nodes per thread + a 4y x 8z strip), 50^3 exact."""
    lib = M.load()
    for n, bound in ((30, 1.03), (50, 1.0001), (40, 1.25), (200, 1.12)):
        dims = (n, n, 12 if n == 200 else n)
        t = tiling(lib, dims, sphere_types(*dims))
        executed = t["ntiles"] * t["cb"] * 32 * t["tn"]
        assert executed <= bound * (dims[0] - 2) * (dims[1] - 2) * (dims[2] - 2) * (1.35 if n == 200 else 1.0), (n, executed)
