// Tile-level phases of the batched r2r transform kernels.  Each phase is a
// function of (tid, nthr) only, so a CUDA block runs
//     phase(tid); __syncthreads(); phase(tid); ...
// and the CPU emulator (tests/emu/emu_fft.cpp) runs every phase for
// tid = 0..nthr-1 in turn -- same code, same index arithmetic.
#pragma once
#include "fft_engine.cuh"

namespace cb {

// geometry of one launch
template <class T> struct FftArgs {
  FftDev<T> P;
  const T* in;
  T* out;
  long long in_es, out_es;  // element stride along the transform direction
  long long in_ls, out_ls;  // stride between consecutive lines of a group
  long long in_gs, out_gs;  // stride between groups (z planes)
  int lines_per_group;      // x transforms: ny ; y transforms: nx
  int ngroups;              // nz
  int line_len;             // points per line (>= P.n; the tail is copied, not transformed)
  int tile_lines;           // lines per tile
  int ymode;                // 0: lanes along the line, 1: lanes across lines
};

struct Tile {
  long long first_line;  // xmode: flattened (group, line) index of line 0
  int group, l0;         // ymode: group and first line within the group
  int nl;                // live lines in this tile
};

template <class T> CB_HD long long num_tiles(const FftArgs<T>& A) {
  if (A.ymode) {
    long long tpg = (A.lines_per_group + A.tile_lines - 1) / A.tile_lines;
    return tpg * A.ngroups;
  }
  long long tot = (long long)A.lines_per_group * A.ngroups;
  return (tot + A.tile_lines - 1) / A.tile_lines;
}

template <class T> CB_HD Tile make_tile(const FftArgs<T>& A, long long t) {
  Tile tl;
  if (A.ymode) {
    int tpg = (A.lines_per_group + A.tile_lines - 1) / A.tile_lines;
    tl.group = (int)(t / tpg);
    tl.l0 = (int)(t - (long long)tl.group * tpg) * A.tile_lines;
    int rem = A.lines_per_group - tl.l0;
    tl.nl = rem < A.tile_lines ? rem : A.tile_lines;
    tl.first_line = 0;
  } else {
    long long tot = (long long)A.lines_per_group * A.ngroups;
    tl.first_line = t * A.tile_lines;
    long long rem = tot - tl.first_line;
    tl.nl = rem < A.tile_lines ? (int)rem : A.tile_lines;
    tl.group = 0;
    tl.l0 = 0;
  }
  return tl;
}

template <class T> CB_HD void line_offsets(const FftArgs<T>& A, const Tile& tl, int c, long long& oin, long long& oout) {
  int g, l;
  if (A.ymode) {
    g = tl.group;
    l = tl.l0 + c;
  } else {
    long long gl = tl.first_line + c;
    g = (int)(gl / A.lines_per_group);
    l = (int)(gl - (long long)g * A.lines_per_group);
  }
  oin = (long long)g * A.in_gs + (long long)l * A.in_ls;
  oout = (long long)g * A.out_gs + (long long)l * A.out_ls;
}

// iteration helper: work item = (line c, sub-index b), flattened so that no
// thread idles when a phase has fewer items per line than threads.
//   xmode: w = c*nsub + b (b fastest across lanes: coalesced along the line)
//   ymode: w = b*CX + c   (c fastest across lanes: coalesced across lines)
struct It {
  int c, b, dc, db, nsub, nl, y;
  CB_HD bool ok() const { return y ? (b < nsub) : (c < nl); }
  CB_HD void next() {
    b += db;
    if (!y) {
      c += dc;
      if (b >= nsub) { b -= nsub; ++c; }
    }
  }
};
template <class T> CB_HD It item_begin(const FftArgs<T>& A, int nl, int nsub, int tid, int nthr) {
  It it;
  it.nsub = nsub; it.nl = nl; it.y = A.ymode;
  if (A.ymode) {
    it.c = tid % A.tile_lines;
    it.b = tid / A.tile_lines;
    it.db = nthr / A.tile_lines;
    it.dc = 0;
    if (it.c >= nl) it.b = nsub;
  } else {
    it.c = tid / nsub;
    it.b = tid - it.c * nsub;
    it.dc = nthr / nsub;
    it.db = nthr - it.dc * nsub;
  }
  return it;
}

// per-line global offsets, recomputed only when the line changes
template <class T> struct LineCache {
  int c = -1;
  long long oin = 0, oout = 0;
  CB_HD void seek(const FftArgs<T>& A, const Tile& tl, int cc) {
    if (cc != c) { c = cc; line_offsets(A, tl, cc, oin, oout); }
  }
};

// ---------------------------------------------------------------- forward ---
template <class T, class Lay>
CB_HD void phase_fwd_load(const FftArgs<T>& A, const Tile& tl, T* s, const Lay& lay, int tid, int nthr) {
  const int n = A.P.n;
  LineCache<T> lc;
  const bool copy_tail = (A.in != A.out) && (A.line_len > n);
  for (It it = item_begin(A, tl.nl, A.line_len, tid, nthr); it.ok(); it.next()) {
    lc.seek(A, tl, it.c);
    const T v = A.in[lc.oin + (long long)it.b * A.in_es];
    if (it.b < n) fwd_load_item(A.P, s, lay, it.c, it.b, v);
    else if (copy_tail) A.out[lc.oout + (long long)it.b * A.out_es] = v;
  }
}

template <class T, class Lay>
CB_HD void phase_stage(const FftArgs<T>& A, const Tile& tl, T* s, const Lay& lay, int st, int tid, int nthr) {
  const int M = A.P.M;
  int Ns = M;
  for (int q = 0; q < st; ++q) Ns /= A.P.radix[q];
  const int R = A.P.radix[st];
  const int Bs = M / Ns;
  for (It it = item_begin(A, tl.nl, M / R, tid, nthr); it.ok(); it.next())
    stage_item_dyn<T, Lay>(R, s, lay, it.c, it.b, Ns, Bs, A.P.tw, M);
}

template <class T> struct OutStore {
  T* dst;
  long long es;
  CB_HD void operator()(int idx, T v) const { dst[(long long)idx * es] = v; }
};
template <class T> struct InFetch {
  const T* src;
  long long es;
  CB_HD T operator()(int idx) const { return src[(long long)idx * es]; }
};

template <class T, class Lay>
CB_HD void phase_fwd_post(const FftArgs<T>& A, const Tile& tl, const T* s, const Lay& lay, int tid, int nthr) {
  LineCache<T> lc;
  for (It it = item_begin(A, tl.nl, A.P.M / 2 + 1, tid, nthr); it.ok(); it.next()) {
    lc.seek(A, tl, it.c);
    OutStore<T> out{A.out + lc.oout, A.out_es};
    fwd_post_item(A.P, s, lay, it.c, it.b, out);
  }
}

// --------------------------------------------------------------- backward ---
template <class T, class Lay>
CB_HD void phase_bwd_pre(const FftArgs<T>& A, const Tile& tl, T* s, const Lay& lay, int tid, int nthr) {
  LineCache<T> lc;
  for (It it = item_begin(A, tl.nl, A.P.M / 2 + 1, tid, nthr); it.ok(); it.next()) {
    lc.seek(A, tl, it.c);
    InFetch<T> in{A.in + lc.oin, A.in_es};
    bwd_pre_item(A.P, s, lay, it.c, it.b, in);
  }
  if ((A.in != A.out) && (A.line_len > A.P.n)) {
    const int ntail = A.line_len - A.P.n;
    for (It it = item_begin(A, tl.nl, ntail, tid, nthr); it.ok(); it.next()) {
      lc.seek(A, tl, it.c);
      const long long i = A.P.n + it.b;
      A.out[lc.oout + i * A.out_es] = A.in[lc.oin + i * A.in_es];
    }
  }
}

template <class T, class Lay>
CB_HD void phase_bwd_out(const FftArgs<T>& A, const Tile& tl, const T* s, const Lay& lay, int tid, int nthr) {
  LineCache<T> lc;
  for (It it = item_begin(A, tl.nl, A.P.n, tid, nthr); it.ok(); it.next()) {
    lc.seek(A, tl, it.c);
    A.out[lc.oout + (long long)it.b * A.out_es] = bwd_out_item(A.P, s, lay, it.c, it.b);
  }
}

// shared-memory elements (of T) a tile needs
template <class T> CB_HD size_t tile_smem_elems(const FftArgs<T>& A) {
  if (A.ymode) return (size_t)A.P.n * A.tile_lines;
  return (size_t)LayX::line_len(A.P.M) * A.tile_lines;
}

}  // namespace cb
