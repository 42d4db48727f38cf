// Compile-time tiling of the lower triangle of a DP×DP Gram matrix into 8×8 DMMA tiles and their assignment to
// the warps of a CTA. Everything here is evaluated by the front end (constant expressions only); nothing lands in
// device memory.
#pragma once
#include <cstdint>
#include <utility>

namespace bdf {

// Tile order: super-rows of 3 block-rows, inside a super-row block-column-major. A run of consecutive tiles then
// touches few distinct 8-wide blocks, so a warp that owns a run needs few operand fragments per k-step.
template <int NB>
struct TileOrder {
  static constexpr int NT = NB * (NB + 1) / 2;
  int I[NT > 0 ? NT : 1];
  int J[NT > 0 ? NT : 1];
  constexpr TileOrder() : I{}, J{} {
    int t = 0;
    for (int a = 0; a < NB; a += 3)
      for (int j = 0; j < NB; j++)
        for (int i = a; i < a + 3 && i < NB; i++)
          if (j <= i) {
            I[t] = i;
            J[t] = j;
            t++;
          }
  }
};

template <int DP_, int NW_>
struct TileCfg {
  static constexpr int DP = DP_;
  static constexpr int NW = NW_;
  static constexpr int NB = DP / 8;
  static constexpr int NT = NB * (NB + 1) / 2;
  static constexpr int TPW = (NT + NW - 1) / NW;  // max tiles per warp
  static constexpr TileOrder<NB> ord{};

  static constexpr int t0(int w) { return w * TPW < NT ? w * TPW : NT; }
  static constexpr int t1(int w) { return (w + 1) * TPW < NT ? (w + 1) * TPW : NT; }
  static constexpr int ntiles(int w) { return t1(w) - t0(w); }
  static constexpr int tileI(int w, int t) { return ord.I[t0(w) + t]; }
  static constexpr int tileJ(int w, int t) { return ord.J[t0(w) + t]; }
  // bitmask of the 8-wide blocks whose fragments warp w needs
  static constexpr uint32_t mask(int w) {
    uint32_t m = 0;
    for (int t = t0(w); t < t1(w); t++) m |= (1u << ord.I[t]) | (1u << ord.J[t]);
    return m;
  }
  static constexpr int popc(uint32_t m) {
    int c = 0;
    for (; m; m &= m - 1) c++;
    return c;
  }
  static constexpr int nfrag(int w) { return popc(mask(w)); }
  static constexpr int rank_of(int w, int blk) { return popc(mask(w) & ((1u << blk) - 1u)); }
  static constexpr int blk_of(int w, int r) {
    uint32_t m = mask(w);
    for (int b = 0; b < 32; b++)
      if (m & (1u << b)) {
        if (r == 0) return b;
        r--;
      }
    return -1;
  }
  static constexpr int max_nfrag() {
    int m = 0;
    for (int w = 0; w < NW; w++) m = nfrag(w) > m ? nfrag(w) : m;
    return m;
  }
};

// per-(warp, tile) compile-time constants
template <class C, int W, int T>
struct TI {
  static constexpr int I = C::tileI(W, T);
  static constexpr int J = C::tileJ(W, T);
  static constexpr int fa = C::rank_of(W, I);
  static constexpr int fb = C::rank_of(W, J);
};
template <class C, int W, int R>
struct FI {
  static constexpr int blk = C::blk_of(W, R);
};

}  // namespace bdf
