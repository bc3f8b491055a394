// backtrack.cu -- DynamicProgram<float>::argmin (reference src/DynamicProgram.cpp:190-255): walk from each
// root hit down the part tree through the back-pointers.  The DP (dt.cu) stores, per (component, part),
//   ik[pm]     best child mixture for parent mixture pm at the parent's cell,
//   ixdt[mm]   row-pass argmax of child mixture mm (stored [x][y]),   iyraw[mm]  column-pass argmax ([y][x]),
// and the reference's Ix/Iy/Ik Mats are recovered on the fly:
//   Ik[p][pm](y,x) = k = ik[pm](y,x)
//   Ix[p][pm](y,x) = ixdt[k](y,x)
//   Iy[p][pm](y,x) = iyraw[k](y, Ix)              (reference composition, include/DistanceTransform.hpp:232-244)
// or, with backptr_mode = 1, the true 2-D argmax  y* = iyraw[k](y,x), x* = ixdt[k](y*,x).
#include "kernels.cuh"

namespace pbd {
namespace {

__global__ void __launch_bounds__(64)
backtrack(const Geometry* __restrict__ g, const int* __restrict__ parent, const int* __restrict__ nparts,
          const int* __restrict__ cm_slot, const int* __restrict__ pm_slot, const unsigned short* __restrict__ ixdt,
          const unsigned short* __restrict__ iyraw, const unsigned char* __restrict__ ik, const unsigned char* __restrict__ rooti,
          int ncomp, int ncm, int npm, const Hit* __restrict__ hits, const int* __restrict__ nhits, int max_hits, int mode,
          int out_parts, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(*nhits, max_hits);
  if (i >= n) return;
  const Hit h = hits[i];
  const LevelDesc& L = g->lv[h.level];
  const size_t ct = (size_t)g->cells_total;
  const int np = nparts[h.comp];
  int* xs = out + (size_t)i * 3 * out_parts;      // [hit][x|y|m][out_parts]
  int* ys = xs + out_parts;
  int* ms = ys + out_parts;
  xs[0] = h.x; ys[0] = h.y;
  ms[0] = rooti[((size_t)h.frame * ncomp + h.comp) * ct + L.cell_off + (size_t)h.y * L.ow + h.x];
  for (int p = 1; p < np; ++p) {                                        // :219-235
    const int par = parent[h.comp * kMaxParts + p];
    const int x = xs[par], y = ys[par], m = ms[par];
    const size_t cell = (size_t)L.cell_off + (size_t)y * L.ow + x;
    const int k = ik[((size_t)h.frame * npm + pm_slot[(h.comp * kMaxParts + p) * kMaxMix + m]) * ct + cell];
    const size_t cmb = ((size_t)h.frame * ncm + cm_slot[(h.comp * kMaxParts + p) * kMaxMix + k]) * ct;
    int xp, yp;
    if (mode == 0) {                                                   // ixdt is stored transposed: [x][y]
      xp = ixdt[cmb + L.cell_off + (size_t)x * L.oh + y];
      yp = iyraw[cmb + L.cell_off + (size_t)y * L.ow + xp];
    } else {
      yp = iyraw[cmb + cell];
      xp = ixdt[cmb + L.cell_off + (size_t)x * L.oh + yp];
    }
    xs[p] = xp; ys[p] = yp; ms[p] = k;
  }
}

__global__ void __launch_bounds__(256)
expand_backptr(int cell_off, int oh, int ow, size_t ct, int frame, int ncm, int npm, const int* __restrict__ cm_slots,
               int pm_slot, const unsigned short* __restrict__ ixdt, const unsigned short* __restrict__ iyraw,
               const unsigned char* __restrict__ ik, int mode, int* __restrict__ ix, int* __restrict__ iy, int* __restrict__ ikout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= oh * ow) return;
  const int x = idx % ow, y = idx / ow;
  const size_t cell = (size_t)cell_off + idx;
  const int k = ik[((size_t)frame * npm + pm_slot) * ct + cell];
  const size_t cmb = ((size_t)frame * ncm + cm_slots[k]) * ct;
  int xp, yp;
  if (mode == 0) {                                                     // ixdt is stored transposed: [x][y]
    xp = ixdt[cmb + cell_off + (size_t)x * oh + y];
    yp = iyraw[cmb + cell_off + (size_t)y * ow + xp];
  } else {
    yp = iyraw[cmb + cell];
    xp = ixdt[cmb + cell_off + (size_t)x * oh + yp];
  }
  ix[idx] = xp; iy[idx] = yp; ikout[idx] = k;
}

}  // namespace

int launch_backtrack(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const BacktrackTables& t, int ncomp, int ncm, int npm,
                     const Hit* d_hits, const int* d_nhits, int max_hits, int backptr_mode, int out_parts, int* d_out_xym, cudaStream_t s) {
  (void)g;
  if (max_hits <= 0) return 0;
  // the hit count lives on the device; launch for the capacity and let surplus threads exit
  backtrack<<<(max_hits + 63) / 64, 64, 0, s>>>(d_g, t.parent, t.nparts, t.cm_slot, t.pm_slot, b.ixdt, b.iyraw, b.ik, b.rooti, ncomp, ncm,
                                                npm, d_hits, d_nhits, max_hits, backptr_mode, out_parts, d_out_xym);
  return 1;
}

int launch_expand_backptr(const Geometry& g, const DeviceBuffers& b, int frame, int level, int ncm, int npm, const int* d_cm_slots,
                          int pm_slot, int backptr_mode, int* d_ix, int* d_iy, int* d_ik, cudaStream_t s) {
  const LevelDesc& L = g.lv[level];
  const int n = L.oh * L.ow;
  if (n <= 0) return 0;
  expand_backptr<<<(n + 255) / 256, 256, 0, s>>>(L.cell_off, L.oh, L.ow, (size_t)g.cells_total, frame, ncm, npm, d_cm_slots, pm_slot, b.ixdt,
                                                 b.iyraw, b.ik, backptr_mode, d_ix, d_iy, d_ik);
  return 1;
}

}  // namespace pbd
