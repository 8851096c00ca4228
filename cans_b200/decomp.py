"""z-slab pencil decomposition over the GPUs of one box -- host-side index arithmetic.

Mirrors the split rule of the reference's decomposition libraries (first `n mod P` ranks get one
extra point: /root/reference/dependencies/2decomp-fft/src/decomp_2d.f90:1018-1029,
dependencies/cuDecomp/src/cudecomp.cc:1348-1357) and the pencil shapes of `initmpi`
(src/initmpi.f90:198-260) for `ipencil_axis = 1`, `dims = [1, P]`:

    x pencil of rank r (what main.f90 sees):     (nx, ny, nz_r),   z in [zs[r], zs[r+1])
    z pencil of rank r (what gaussel sees):      (nx, ny_r, nz),   y in [ys[r], ys[r+1])

The same arithmetic lives in capi.cu (`split_starts`, `build_dist_tables`); the GPU tests check
that both agree, the CPU (gloo) tests run the exchange these tables describe with the oracle.
"""
from __future__ import annotations

from typing import List


def split_starts(n: int, nranks: int) -> List[int]:
    base, rem = divmod(n, nranks)
    st = [0]
    for r in range(nranks):
        st.append(st[-1] + base + (1 if r < rem else 0))
    return st


class SlabDecomp:
    def __init__(self, ng, nranks: int, rank: int):
        if nranks < 1 or not 0 <= rank < nranks:
            raise ValueError("bad rank / nranks")
        if nranks > ng[1] or nranks > ng[2]:
            raise ValueError("more ranks than y or z planes")
        self.ng = [int(v) for v in ng]
        self.nranks, self.rank = nranks, rank
        self.ys = split_starts(self.ng[1], nranks)
        self.zs = split_starts(self.ng[2], nranks)

    # extents as initmpi returns them (1-based lo)
    @property
    def n(self):
        return [self.ng[0], self.ng[1], self.zs[self.rank + 1] - self.zs[self.rank]]

    @property
    def lo(self):
        return [1, 1, self.zs[self.rank] + 1]

    @property
    def n_z(self):
        return [self.ng[0], self.ys[self.rank + 1] - self.ys[self.rank], self.ng[2]]

    @property
    def lo_z(self):
        return [1, self.ys[self.rank] + 1, 1]

    def y_range(self, r=None):
        r = self.rank if r is None else r
        return self.ys[r], self.ys[r + 1]

    def z_range(self, r=None):
        r = self.rank if r is None else r
        return self.zs[r], self.zs[r + 1]

    # ---- what travels in the two exchanges; arrays are [k, j, i]
    def forward_blocks(self, slab):
        """y-transformed slab (nz_r, ny, nx) -> list over destination ranks s of slab[:, ys[s]:ys[s+1], :]"""
        return [slab[:, self.ys[s]:self.ys[s + 1], :] for s in range(self.nranks)]

    def assemble_zpencil(self, blocks):
        """blocks[s] = (nz_s, ny_r, nx) received from rank s -> z pencil (nz, ny_r, nx)"""
        import numpy as np
        return np.concatenate(blocks, axis=0)

    def backward_blocks(self, zpencil):
        """z pencil (nz, ny_r, nx) -> list over destination ranks s of zpencil[zs[s]:zs[s+1]]"""
        return [zpencil[self.zs[s]:self.zs[s + 1]] for s in range(self.nranks)]

    def assemble_slab(self, blocks):
        """blocks[s] = (nz_r, ny_s, nx) received from rank s -> slab (nz_r, ny, nx)"""
        import numpy as np
        return np.concatenate(blocks, axis=1)
