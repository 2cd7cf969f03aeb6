// Monte Carlo BLOCK SWEEP for large lattices: one CTA owns one tile of the brick-ordered supercell for a whole sweep.
//
// mc_colour_kernel (asd_mc.cuh) needs one launch per atom colour, and every launch gathers the 50 neighbours of its
// atoms from L2 / HBM again: 2 GB of DRAM traffic per sweep of bcc Fe 128^3 for 1.07 GB of algorithmic bytes, bound by
// the L2 gather rate (profiles/r01: 23 % of the 256 B/attempt roofline).  Here the TILES are coloured as well (two
// tiles may run concurrently iff no atom of one has a neighbour in the other; greedy colouring of that tile graph,
// 8 classes for a periodic 3-D grid of tiles), and one launch per tile colour does, per tile:
//   (A) stage emomM of the tile's gather list in shared memory ONCE (the same lists the LLG kernels use, but ordered
//       with the x-residue split of asd_tiles.cuh so that the atoms of one atom colour are consecutive positions);
//   (B) draw the trial moves (Metropolis) / the polar draws (heat bath) of all atoms of the tile, all threads busy --
//       mc_evolve also draws every trial move before it sweeps (montecarlo.f90:131-160);
//   (C) sweep the atom colours one after the other: the atoms of a colour are updated concurrently, two lanes per atom
//       (each sums half of the neighbour list from shared memory, one warp shuffle joins them), the new moment goes
//       back into the shared-memory copy and into `cur`; __syncthreads() between colours.
// Every update sees exactly the state a sequential sweep in the order (tile colour, tile, atom colour, atom) would
// see: the chain is a sequential single-site chain in that order (asd_get_mc_visit_order reports it; the tests replay
// it through a CPU restatement of mc_evolve with the same draws).  DRAM traffic per sweep: gather lists (3.1 spins
// per atom) + one write per accepted move + the 16-bit position words.
//
// Restated pieces (same formulas as mc_update_site, asd_mc.cuh):
//   trial move   choose_random_flip   source/MonteCarlo/montecarlo_common.f90:25-79
//   delta E      calculate_energy     source/MonteCarlo/montecarlo_common.f90:431-865
//   Metropolis   flip_a               source/MonteCarlo/montecarlo_common.f90:190-200
//   heat bath    flip_h               source/MonteCarlo/montecarlo_common.f90:371-422
#pragma once
#include "asd_mc.cuh"

namespace asd {

struct McBlock {
   int ts;                   // slots per tile (256 or 1024)
   int ucap;                 // row stride of ulist
   int ncol;                 // atom colours
   const int* __restrict__ ulist;              // [ntile][ucap] gather lists (x-residue split order)
   const int* __restrict__ ucount;             // [ntile]
   const uint4* __restrict__ nl16;             // [ceil(z/8)][Npad] exchange neighbours as positions in the tile's list
   const uint4* __restrict__ dm16;             // [ceil(zdm/8)][Npad] or null
   const uint4* __restrict__ bq16;             // [ceil(zbq/8)][Npad] or null
   const unsigned short* __restrict__ selfpos; // [Npad] position of the atom itself in its tile's list
   const unsigned short* __restrict__ corder;  // [ntile][ts] local slots of the tile's real atoms sorted by (colour, slot)
   const int* __restrict__ cstart;             // [ntile][ncol + 1] first corder index of every colour
   const int* __restrict__ tilelist;           // tiles sorted by tile colour
};

// per tile: local slots sorted by (colour, slot) and the start of every colour class
__global__ void __launch_bounds__(64)
mc_block_order_kernel(int Nown, int ts, int ncol, const unsigned char* __restrict__ col, unsigned short* __restrict__ corder,
                      int* __restrict__ cstart) {
   __shared__ int cnt[256], start[257];
   const int tile = blockIdx.x;
   const int s0 = tile * ts, s1 = min(s0 + ts, Nown);
   for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
      int n = 0;
      for (int s = s0; s < s1; s++) n += (col[s] == (unsigned char)c);
      cnt[c] = n;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      int a = 0;
      for (int c = 0; c < ncol; c++) { start[c] = a; a += cnt[c]; }
      start[ncol] = a;
   }
   __syncthreads();
   for (int c = threadIdx.x; c <= ncol; c += blockDim.x) cstart[(size_t)tile * (ncol + 1) + c] = start[c];
   for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
      int n = start[c];
      for (int s = s0; s < s1; s++)
         if (col[s] == (unsigned char)c) corder[(size_t)tile * ts + n++] = (unsigned short)(s - s0);
   }
   for (int q = start[ncol] + threadIdx.x; q < ts; q += blockDim.x) corder[(size_t)tile * ts + q] = 0xffffu;
}

// tiles a tile gathers from: adj[tile][cap] (unsorted), nadj[tile] (cap + 1 when the list overflowed)
__global__ void __launch_bounds__(256)
mc_block_adjacency_kernel(int ntile, int ts, int ucap, const int* __restrict__ ulist, const int* __restrict__ ucount, int cap,
                          int* __restrict__ adj, int* __restrict__ nadj) {
   extern __shared__ unsigned int bits[];
   __shared__ int n;
   const int tile = blockIdx.x, nw = (ntile + 31) / 32;
   for (int q = threadIdx.x; q < nw; q += blockDim.x) bits[q] = 0u;
   if (threadIdx.x == 0) n = 0;
   __syncthreads();
   const int cnt = ucount[tile];
   for (int u = threadIdx.x; u < cnt; u += blockDim.x) {
      const int other = ulist[(size_t)tile * ucap + u] / ts;
      if (other != tile && other < ntile) atomicOr(&bits[other >> 5], 1u << (other & 31));
   }
   __syncthreads();
   for (int q = threadIdx.x; q < nw; q += blockDim.x) {
      unsigned int w = bits[q];
      while (w) {
         const int b = __ffs(w) - 1;
         w &= w - 1;
         const int at = atomicAdd(&n, 1);
         if (at < cap) adj[(size_t)tile * cap + at] = q * 32 + b;
      }
   }
   __syncthreads();
   if (threadIdx.x == 0) nadj[tile] = min(n, cap + 1);
}

// anisotropy ENERGY of the current moment c and the trial moment tr (calculate_energy, montecarlo_common.f90:536-566)
template <bool REDUCED>
__device__ __forceinline__ void aniso_energy(const Tables& t, int i, int ih, const double c[3], const double tr[3], double& e_c,
                                             double& e_t) {
   if (!t.do_aniso) return;
   const bool rows = REDUCED && t.aniso_rows;
   const int ta = rows ? (int)t.aniso_small[ih][0] : __ldg(t.taniso + i);
   if (ta != 1 && ta != 2 && ta != 7) return;
   const double k1 = rows ? t.aniso_small[ih][1] : __ldg(t.kaniso + i), k2 = rows ? t.aniso_small[ih][2] : __ldg(t.kaniso + t.Npad + i);
   if (ta == 1 || ta == 7) {
      const double ex = rows ? t.aniso_small[ih][3] : __ldg(t.eaniso + i), ey = rows ? t.aniso_small[ih][4] : __ldg(t.eaniso + t.Npad + i),
                   ez = rows ? t.aniso_small[ih][5] : __ldg(t.eaniso + 2 * (size_t)t.Npad + i);
      const double tta = c[0] * ex + c[1] * ey + c[2] * ez;
      const double ttb = tr[0] * ex + tr[1] * ey + tr[2] * ez;
      e_c += k1 * (tta * tta) + k2 * (tta * tta) * (tta * tta);
      e_t += k1 * (ttb * ttb) + k2 * (ttb * ttb) * (ttb * ttb);
   }
   if (ta == 2 || ta == 7) {
      const double c4 = (c[0] * c[0]) * (c[1] * c[1]) + (c[1] * c[1]) * (c[2] * c[2]) + (c[2] * c[2]) * (c[0] * c[0]);
      const double c6 = (c[0] * c[0]) * (c[1] * c[1]) * (c[2] * c[2]);
      const double t4 = (tr[0] * tr[0]) * (tr[1] * tr[1]) + (tr[1] * tr[1]) * (tr[2] * tr[2]) + (tr[2] * tr[2]) * (tr[0] * tr[0]);
      const double t6 = (tr[0] * tr[0]) * (tr[1] * tr[1]) * (tr[2] * tr[2]);
      if (ta == 2) { e_c += -k1 * c4 - k2 * c6; e_t += -k1 * t4 - k2 * t6; }
      else {
         const double s = rows ? t.aniso_small[ih][6] : __ldg(t.sb + i);
         e_c += (k1 * s) * c4 + (k2 * s) * c6;
         e_t += (k1 * s) * t4 + (k2 * s) * t6;
      }
   }
}

__device__ __forceinline__ void unpack16(const uint4& w, unsigned li[8]) {
   li[0] = w.x & 0xffffu; li[1] = w.x >> 16; li[2] = w.y & 0xffffu; li[3] = w.y >> 16;
   li[4] = w.z & 0xffffu; li[5] = w.z >> 16; li[6] = w.w & 0xffffu; li[7] = w.w >> 16;
}

// APT: atoms per thread of the draw phase = tile slots / 256.  XS: the layout has DM and / or BQ tables.  HB: heat bath.
// Reduced Hamiltonians only (couplings [NH][z] staged in shared memory).
template <int APT, bool XS, bool HB>
__global__ void __launch_bounds__(256, APT == 4 ? 2 : 4)
mc_block_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, const __grid_constant__ McBlock mb,
                const int class_first, SpinVec* __restrict__ cur) {
   constexpr int TS = 256 * APT, SB = 6;
   constexpr unsigned FULL = 0xffffffffu;
   extern __shared__ double sm[];
   const int tid = threadIdx.x;
   const int tile = __ldg(mb.tilelist + class_first + (int)blockIdx.x), k = blockIdx.y;
   SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
   const size_t Npad = t.Npad;
   // shared memory: couplings (exchange | DM | BQ) | draws[TS][4] | emomM of the gather list [ucap][3]
   const int ncp = t.NH * t.z, ndm = XS ? t.NH * t.zdm * 3 : 0, nbq = XS ? t.NH * t.zbq : 0;
   double* __restrict__ scp = sm;
   double* __restrict__ sdm = scp + ncp;
   double* __restrict__ sbq = sdm + ndm;
   double* __restrict__ draws = sm + ((ncp + ndm + nbq + 3) & ~3);
   double* __restrict__ s3 = draws + 4 * TS;
   for (int q = tid; q < ncp; q += 256) scp[q] = __ldg(t.cp + q);
   if (XS) {
      for (int q = tid; q < ndm; q += 256) sdm[q] = __ldg(t.dmv + q);
      for (int q = tid; q < nbq; q += 256) sbq[q] = __ldg(t.jbq + q);
   }
   // ---- (A) gather list -> shared memory ----
   const int cnt = __ldg(mb.ucount + tile);
   const int* __restrict__ ul = mb.ulist + (size_t)tile * mb.ucap;
   for (int u0 = tid; u0 < cnt; u0 += SB * 256) {
      int sl[SB];
#pragma unroll
      for (int a = 0; a < SB; a++) sl[a] = (u0 + a * 256 < cnt) ? __ldg(ul + u0 + a * 256) : 0;
      SpinVec v[SB];
#pragma unroll
      for (int a = 0; a < SB; a++) v[a] = S[sl[a]];
#pragma unroll
      for (int a = 0; a < SB; a++)
         if (u0 + a * 256 < cnt) {
            double* __restrict__ m = s3 + 3 * (u0 + a * 256);
            m[0] = v[a].x * v[a].m; m[1] = v[a].y * v[a].m; m[2] = v[a].z * v[a].m;
         }
   }
   // ---- (B) draws of every atom of the tile, in colour order (position `pos` of corder) ----
   const int* __restrict__ cs = mb.cstart + (size_t)tile * (mb.ncol + 1);
   const unsigned short* __restrict__ co = mb.corder + (size_t)tile * TS;
   const int nreal = __ldg(cs + mb.ncol);
   const double pi = 3.141592653589793;
#pragma unroll 1
   for (int a = 0; a < APT; a++) {
      const int pos = a * 256 + tid;
      if (pos < nreal) {
         const int i = tile * TS + (int)co[pos];
         const int o = __ldg(t.orig + i);
         double u[4];
         uniform4(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 1u, u);
         double d0, d1, d2, d3;
         if (HB) {
            // flip_h: u[0] picks cos(theta) about the local field (needs the field: phase C), u[1] the azimuth
            double sphi, cphi;
            sincos(pi * (2.0 * u[1] - 1.0), &sphi, &cphi);
            d0 = u[0]; d1 = cphi; d2 = sphi; d3 = 0.0;
         } else {
            // choose_random_flip: the moment has not been visited in this sweep yet, S[i] is its value at the sweep start
            const SpinVec own = S[i];
            const int ftype = (int)floor(3.0 * u[0]);
            if (ftype == 0) {
               double sphi, cphi;
               sincos(u[1] * 2 * pi, &sphi, &cphi);
               const double ct = 1.0 - 2.0 * u[2];
               const double st = sqrt(fmax(1.0 - ct * ct, 0.0));
               d0 = st * cphi; d1 = st * sphi; d2 = ct;
            } else if (ftype == 1) {
               double g0, g1, g2;
               gauss3f(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 2u, g0, g1, g2);
               const double ax = own.x + g0 * p.delta, ay = own.y + g1 * p.delta, az = own.z + g2 * p.delta;
               const double l = sqrt(ax * ax + ay * ay + az * az);
               const double rl = 1.0 / l; d0 = ax * rl; d1 = ay * rl; d2 = az * rl;
            } else {
               d0 = -own.x; d1 = -own.y; d2 = -own.z;
            }
            d3 = u[3];
         }
         double* __restrict__ dr = draws + 4 * pos;
         dr[0] = d0; dr[1] = d1; dr[2] = d2; dr[3] = d3;
      }
   }
   __syncthreads();
   // ---- (C) atom colours, one after the other; two lanes (tid, tid ^ 16) share one atom ----
   const int sub = (tid & 15) + 16 * (tid >> 5), half = (tid >> 4) & 1;
   const double beta_h = 1.0 / p.k_bolt / (p.temprescale * p.temperature);
   const double beta_m = 1.0 / p.k_bolt / (p.temprescale * p.temperature + 1.0e-15);
   for (int c = 0; c < mb.ncol; c++) {
      const int n0 = __ldg(cs + c), n1 = __ldg(cs + c + 1);
      for (int a0 = n0; a0 < n1; a0 += 128) {
         const int idx = a0 + sub;
         const bool act = idx < n1;
         int i = 0, ih = 0;
         SpinVec own;
         own.x = 0.0; own.y = 0.0; own.z = 1.0; own.m = 0.0;
         double dr[4] = {0.0, 0.0, 0.0, 0.0};
         double f[3] = {0.0, 0.0, 0.0};      // Heisenberg + DM field of the frozen neighbours (half of the lists per lane)
         double g[3] = {0.0, 0.0, 0.0};      // HB: biquadratic field at the current moment; Metropolis: g[0], g[1] = BQ energies
         double cm[3] = {0.0, 0.0, 0.0}, tm[3] = {0.0, 0.0, 0.0};
         if (act) {
            i = tile * TS + (int)co[idx];
            ih = __ldg(t.ham + i);
            const int n = __ldg(t.lsize + ih);
            // the first position words of this lane's half of the list: issued before anything else of the atom
            uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
            if (8 * half < n) w0 = __ldg(mb.nl16 + (size_t)half * Npad + i);
            if (8 * (half + 2) < n) w1 = __ldg(mb.nl16 + (size_t)(half + 2) * Npad + i);
            own = S[i];
            const double* __restrict__ d = draws + 4 * idx;
            dr[0] = d[0]; dr[1] = d[1]; dr[2] = d[2]; dr[3] = d[3];
            cm[0] = own.x * own.m; cm[1] = own.y * own.m; cm[2] = own.z * own.m;
            if (!HB) { tm[0] = dr[0] * own.m; tm[1] = dr[1] * own.m; tm[2] = dr[2] * own.m; }
            const double* __restrict__ crow = scp + ih * t.z;
            for (int q = half; 8 * q < n; q += 2) {
               const uint4 w = w0;
               w0 = w1;
               if (8 * (q + 4) < n) w1 = __ldg(mb.nl16 + (size_t)(q + 4) * Npad + i);
               unsigned li[8];
               unpack16(w, li);
               if (8 * q + 8 <= n) {
#pragma unroll
                  for (int u = 0; u < 8; u++) {
                     const double* __restrict__ m = s3 + li[u] * 3u;
                     const double cc = crow[8 * q + u];
                     f[0] = fma(cc, m[0], f[0]); f[1] = fma(cc, m[1], f[1]); f[2] = fma(cc, m[2], f[2]);
                  }
               } else {
#pragma unroll
                  for (int u = 0; u < 8; u++)
                     if (8 * q + u < n) {
                        const double* __restrict__ m = s3 + li[u] * 3u;
                        const double cc = crow[8 * q + u];
                        f[0] = fma(cc, m[0], f[0]); f[1] = fma(cc, m[1], f[1]); f[2] = fma(cc, m[2], f[2]);
                     }
               }
            }
            if (XS && t.zdm > 0) {
               const int nd = __ldg(t.dmsize + ih);
               for (int q = half; 8 * q < nd; q += 2) {
                  unsigned li[8];
                  unpack16(__ldg(mb.dm16 + (size_t)q * Npad + i), li);
#pragma unroll
                  for (int u = 0; u < 8; u++)
                     if (8 * q + u < nd) {
                        const double* __restrict__ m = s3 + li[u] * 3u;
                        const double* __restrict__ D = sdm + (ih * t.zdm + 8 * q + u) * 3;
                        dm_term(D[0], D[1], D[2], m[0], m[1], m[2], f[0], f[1], f[2]);
                     }
               }
            }
            if (XS && t.zbq > 0) {
               const int nb = __ldg(t.bqsize + ih);
               for (int q = half; 8 * q < nb; q += 2) {
                  unsigned li[8];
                  unpack16(__ldg(mb.bq16 + (size_t)q * Npad + i), li);
#pragma unroll
                  for (int u = 0; u < 8; u++)
                     if (8 * q + u < nb) {
                        const double* __restrict__ m = s3 + li[u] * 3u;
                        const double jb = sbq[ih * t.zbq + 8 * q + u];
                        if (HB) bq_term(jb, m[0], m[1], m[2], cm[0], cm[1], cm[2], g[0], g[1], g[2]);
                        else {
                           const double dc = m[0] * cm[0] + m[1] * cm[1] + m[2] * cm[2], dt = m[0] * tm[0] + m[1] * tm[1] + m[2] * tm[2];
                           g[0] += jb * dc * dc;
                           g[1] += jb * dt * dt;
                        }
                     }
               }
            }
         }
         __syncwarp();
#pragma unroll
         for (int a = 0; a < 3; a++) f[a] += __shfl_xor_sync(FULL, f[a], 16);
         if (XS) {
#pragma unroll
            for (int a = 0; a < 3; a++) g[a] += __shfl_xor_sync(FULL, g[a], 16);
         }
         if (act && half == 0) {
            const double m = own.m;
            SpinVec out = own;
            bool changed;
            if (HB) {
               // ---- flip_h: total field = beff1 + beff2 of effective_field_single; external field from the tables ----
               double bs[3] = {f[0], f[1], f[2]}, bq[3] = {g[0], g[1], g[2]}, h[3];
               aniso_field<true>(t, i, ih, cm[0], cm[1], cm[2], bs[0], bs[1], bs[2], bq[0], bq[1], bq[2]);
               ext_field(t, i, k, h);
               const double tot[3] = {bs[0] + (bq[0] + h[0]), bs[1] + (bq[1] + h[1]), bs[2] + (bq[2] + h[2])};
               const double zx = beta_h * tot[0] * p.mub * m, zy = beta_h * tot[1] * p.mub * m, zz = beta_h * tot[2] * p.mub * m;
               const double zarg = sqrt(zx * zx + zy * zy + zz * zz);
               const double rzarg = 1.0 / zarg; const double zctheta = zz * rzarg;
               const double zstheta = sqrt(1.0 - zctheta * zctheta) + 1e-14;
               const double rzs = 1.0 / (zarg * zstheta); double zcphi = zx * rzs, zsphi = zy * rzs;
               if (zx == 0.0 && zy == 0.0) { zcphi = 1.0; zsphi = 0.0; }     // degenerate frame (see mc_update_site)
               const double em2 = exp(-2.0 * zarg);
               const double ctheta = 1.0 + rzarg * log((1.0 - em2) * dr[0] + em2 + 1e-14);
               const double stheta = sqrt(fmax(1.0 - ctheta * ctheta, 0.0));
               const double s0 = stheta * dr[1], s1 = stheta * dr[2], s2 = ctheta;
               out.x = zcphi * zctheta * s0 - zsphi * s1 + zcphi * zstheta * s2;
               out.y = zsphi * zctheta * s0 + zcphi * s1 + zsphi * zstheta * s2;
               out.z = -zstheta * s0 + zctheta * s2;
               changed = true;
            } else {
               // ---- calculate_energy + flip_a ----
               double e_c = 0.0, e_t = 0.0;
               aniso_energy<true>(t, i, ih, cm, tm, e_c, e_t);
               e_c -= cm[0] * f[0] + cm[1] * f[1] + cm[2] * f[2];
               e_t -= tm[0] * f[0] + tm[1] * f[1] + tm[2] * f[2];
               if (XS) { e_c -= g[0]; e_t -= g[1]; }
               e_c -= p.extfield[0] * cm[0] + p.extfield[1] * cm[1] + p.extfield[2] * cm[2];
               e_t -= p.extfield[0] * tm[0] + p.extfield[1] * tm[1] + p.extfield[2] * tm[2];
               const double de = p.mub * (e_t - e_c);
               changed = de <= 0.0 || dr[3] < exp(-beta_m * de);
               if (changed) { out.x = dr[0]; out.y = dr[1]; out.z = dr[2]; }
            }
            if (changed) {
               S[i] = out;
               double* __restrict__ rec = s3 + 3 * (int)__ldg(mb.selfpos + i);
               rec[0] = out.x * m; rec[1] = out.y * m; rec[2] = out.z * m;
            }
         }
      }
      __syncthreads();
   }
}

}  // namespace asd
