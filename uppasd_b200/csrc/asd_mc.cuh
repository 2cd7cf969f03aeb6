// Monte Carlo sweeps on the device: Metropolis and heat bath, one kernel launch per colour class.
//
// The reference visits spins one at a time in a random order (mc_evolve, source/MonteCarlo/montecarlo.f90:
// 44-273; OpenMP sweep racy unless one thread).  Here the conflict graph (union of all neighbour tables,
// symmetrised) is coloured once; all spins of one colour are updated concurrently while every spin they
// interact with is frozen, so each launch is an exact single-site update for every one of its spins.
// This is a different but equally valid Markov chain: parity with the reference is on observables.
//
// Restated pieces:
//   trial move   choose_random_flip   source/MonteCarlo/montecarlo_common.f90:25-79 (Hinzke-Nowak mixture)
//   delta E      calculate_energy     source/MonteCarlo/montecarlo_common.f90:431-865
//   Metropolis   flip_a               source/MonteCarlo/montecarlo_common.f90:190-200
//   heat bath    flip_h               source/MonteCarlo/montecarlo_common.f90:371-422
// Deviation (documented in DESIGN.md): the DM term of delta E uses emomM consistently
// (-m_i . (m_j x D)); the reference mixes emom/emomM there (:611-616), identical for |m| = 1.
#pragma once
#include "asd_device.cuh"
#include "asd_lattice.cuh"
#include <cooperative_groups.h>

namespace asd {

struct McParams {
   int mode;  // 'M' or 'H'
   double temperature, temprescale, k_bolt, mub;
   double extfield[3];  // mc_evolve's uniform field argument (Metropolis Zeeman term)
   unsigned long long seed, sweep;
   int first, count;  // device-slot range [first, first+count) of this colour class (colour-major layout)
   int colour;        // colour updated by this launch (lattice layout, mc_tile_kernel)
   double delta;      // cone width of the Gaussian trial move (montecarlo.f90:142), evaluated once on the host
};

#ifndef ASD_MC_MINB
#define ASD_MC_MINB 4     // CTAs per SM the colour kernel is compiled for
#endif
#ifndef ASD_MC_CHUNK
#define ASD_MC_CHUNK 4    // 256-bit gathers in flight per thread
#endif

// One single-spin update of site i (device slot) of ensemble k with every neighbour frozen: returns true and the new
// spin in `out` when the spin changes (heat bath: always; Metropolis: when the trial move is accepted).
// EXCH = false: the Heisenberg sum was already accumulated by the caller (mc_colour_coop_kernel) and arrives in hx[3].
template <bool REDUCED, bool EXCH = true>
__device__ __forceinline__ bool mc_update_site(const Tables& t, const McParams& p, const SpinVec* __restrict__ S, int i, int k, int o,
                                               int ih, const double* smc, const double* smd, const double* smb, SpinVec& out,
                                               const double* hx = nullptr) {
   const SpinVec own = S[i];
   const double m = own.m;
   double bs[3], bq[3];
   if (!EXCH) { bs[0] = hx[0]; bs[1] = hx[1]; bs[2] = hx[2]; }
   // bilinear field from frozen neighbours (exchange + DM [+ uniaxial]); bq = BQ/cubic field at the CURRENT spin
   site_field<REDUCED, EXCH, ASD_MC_CHUNK>(t, S, i, ih, own, smc, smd, smb, bs, bq);
   out = own;
   double u[4];
   uniform4(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 1u, u);
   const double pi = 3.141592653589793;
   if (p.mode == 'H') {
      // ---- heat bath (flip_h): field = beff1 + beff2 from effective_field_single, external field from the
      //      external_field array / uniform vector of the tables (reference quirk, montecarlo.f90:231-237)
      double h[3];
      ext_field(t, i, k, h);
      const double tot[3] = {bs[0] + (bq[0] + h[0]), bs[1] + (bq[1] + h[1]), bs[2] + (bq[2] + h[2])};
      const double beta = 1.0 / p.k_bolt / (p.temprescale * p.temperature);
      const double zx = beta * tot[0] * p.mub * m, zy = beta * tot[1] * p.mub * m, zz = beta * tot[2] * p.mub * m;
      const double zarg = sqrt(zx * zx + zy * zy + zz * zz);
      // (flip_h divides four times -- zz / zarg, zx / (zarg zstheta), zy / (zarg zstheta), 1 / zarg; here: two reciprocals and products,
      //  last-bit differences at most; the same in every Monte Carlo kernel: an FP64 division costs ~15 dependent instructions)
      const double rzarg = 1.0 / zarg; const double zctheta = zz * rzarg;
      const double zstheta = sqrt(1.0 - zctheta * zctheta) + 1e-14;
      const double rzs = 1.0 / (zarg * zstheta); double zcphi = zx * rzs, zsphi = zy * rzs;
      // Deviation from flip_h (documented in DESIGN.md): a field exactly along +-z makes the reference's frame
      // degenerate (zcphi = zsphi = 0 -> the new spin loses its transverse part and is no longer a unit vector);
      // any azimuth is valid there, take phi_field = 0.
      if (zx == 0.0 && zy == 0.0) { zcphi = 1.0; zsphi = 0.0; }
      const double em2 = exp(-2.0 * zarg);
      const double ctheta = 1.0 + rzarg * log((1.0 - em2) * u[0] + em2 + 1e-14);
      const double stheta = sqrt(fmax(1.0 - ctheta * ctheta, 0.0));
      double sphi, cphi;
      sincos(pi * (2.0 * u[1] - 1.0), &sphi, &cphi);
      const double s0 = stheta * cphi, s1 = stheta * sphi, s2 = ctheta;
      out.x = zcphi * zctheta * s0 - zsphi * s1 + zcphi * zstheta * s2;
      out.y = zsphi * zctheta * s0 + zcphi * s1 + zsphi * zstheta * s2;
      out.z = -zstheta * s0 + zctheta * s2;
      return true;
   }
   // ---- Metropolis: trial move (choose_random_flip) ----
   double nx, ny, nz;
   const int ftype = (int)floor(3.0 * u[0]);
   if (ftype == 0) {
      double sphi, cphi;
      sincos(u[1] * 2 * pi, &sphi, &cphi);
      const double ct = 1.0 - 2.0 * u[2];
      const double st = sqrt(fmax(1.0 - ct * ct, 0.0));
      nx = st * cphi; ny = st * sphi; nz = ct;
   } else if (ftype == 1) {
      double g0, g1, g2;
      gauss3f(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 2u, g0, g1, g2);
      const double ax = own.x + g0 * p.delta, ay = own.y + g1 * p.delta, az = own.z + g2 * p.delta;
      const double l = sqrt(ax * ax + ay * ay + az * az);
      const double rl = 1.0 / l; nx = ax * rl; ny = ay * rl; nz = az * rl;
   } else {
      nx = -own.x; ny = -own.y; nz = -own.z;
   }
   // ---- delta E (calculate_energy): bilinear terms via the frozen-neighbour field, the rest explicitly ----
   const double cx = own.x * m, cy = own.y * m, cz = own.z * m;  // current emomM
   const double tx = nx * m, ty = ny * m, tz = nz * m;           // trial moment
   // exchange + DM:  e = -m_i . f   (f = sum_j J m_j + sum_j m_j x D)
   double fx = bs[0], fy = bs[1], fz = bs[2];
   double e_c = 0.0, e_t = 0.0;
   if (t.do_aniso) {
      // remove the uniaxial field that site_field folded into bs, then add the reference's anisotropy ENERGY
      const int ta = __ldg(t.taniso + i);
      if (ta == 1 || ta == 2 || ta == 7) {
         const double k1 = __ldg(t.kaniso + i), k2 = __ldg(t.kaniso + t.Npad + i);
         if (ta == 1 || ta == 7) {
            const double ex = __ldg(t.eaniso + i), ey = __ldg(t.eaniso + t.Npad + i), ez = __ldg(t.eaniso + 2 * (size_t)t.Npad + i);
            const double tt1 = cx * ex + cy * ey + cz * ez;
            const double tt3 = 2.0 * tt1 * (k1 + 2.0 * k2 * (1.0 - tt1 * tt1));
            fx += tt3 * ex; fy += tt3 * ey; fz += tt3 * ez;
            const double ttb = tx * ex + ty * ey + tz * ez;
            e_c += k1 * (tt1 * tt1) + k2 * (tt1 * tt1) * (tt1 * tt1);
            e_t += k1 * (ttb * ttb) + k2 * (ttb * ttb) * (ttb * ttb);
         }
         if (ta == 2 || ta == 7) {
            const double c4 = (cx * cx) * (cy * cy) + (cy * cy) * (cz * cz) + (cz * cz) * (cx * cx);
            const double c6 = (cx * cx) * (cy * cy) * (cz * cz);
            const double t4 = (tx * tx) * (ty * ty) + (ty * ty) * (tz * tz) + (tz * tz) * (tx * tx);
            const double t6 = (tx * tx) * (ty * ty) * (tz * tz);
            if (ta == 2) {
               // cubic field of taniso 2 was folded into bs as well: remove it
               const double x2 = cx * cx, y2 = cy * cy, z2 = cz * cz;
               fx -= 2.0 * k1 * cx * (y2 + z2) + 2.0 * k2 * cx * (y2 * z2);
               fy -= 2.0 * k1 * cy * (z2 + x2) + 2.0 * k2 * cy * (z2 * x2);
               fz -= 2.0 * k1 * cz * (x2 + y2) + 2.0 * k2 * cz * (x2 * y2);
               e_c += -k1 * c4 - k2 * c6;
               e_t += -k1 * t4 - k2 * t6;
            } else {
               const double s = __ldg(t.sb + i);
               e_c += (k1 * s) * c4 + (k2 * s) * c6;
               e_t += (k1 * s) * t4 + (k2 * s) * t6;
            }
         }
      }
   }
   e_c -= cx * fx + cy * fy + cz * fz;
   e_t -= tx * fx + ty * fy + tz * fz;
   // biquadratic: -j_bq (m_i . m_j)^2 for current and trial moment
   if (t.zbq > 0) {
      const int* __restrict__ nl = t.bql + i;
      const int n = REDUCED ? __ldg(t.bqsize + ih) : t.zbq;
      for (int j = 0; j < n; j++) {
         const int nb = __ldg(nl + (size_t)j * t.Npad);
         const double jb = REDUCED ? (smb ? smb : t.jbq)[(size_t)ih * t.zbq + j] : __ldg(t.jbq + (size_t)j * t.Npad + i);
         const SpinVec v = S[nb];
         const double mx = v.x * v.m, my = v.y * v.m, mz = v.z * v.m;
         const double dc = mx * cx + my * cy + mz * cz, dt = mx * tx + my * ty + mz * tz;
         e_c -= jb * dc * dc;
         e_t -= jb * dt * dt;
      }
   }
   // Zeeman with mc_evolve's extfield argument
   e_c -= p.extfield[0] * cx + p.extfield[1] * cy + p.extfield[2] * cz;
   e_t -= p.extfield[0] * tx + p.extfield[1] * ty + p.extfield[2] * tz;
   const double de = p.mub * (e_t - e_c);
   const double beta = 1.0 / p.k_bolt / (p.temprescale * p.temperature + 1.0e-15);
   if (de <= 0.0 || u[3] < exp(-beta * de)) {
      out.x = nx; out.y = ny; out.z = nz;
      return true;
   }
   return false;
}

// colour-major layout: the slots [first, first+count) are one colour class
template <bool REDUCED>
__global__ void __launch_bounds__(256, ASD_MC_MINB)
mc_colour_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, SpinVec* __restrict__ cur, unsigned int* __restrict__ accepted) {
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   stage_couplings(t, sm, smc, smd, smb);
   const int li = blockIdx.x * blockDim.x + threadIdx.x;
   const int k = blockIdx.y;
   if (li >= p.count) return;
   const int i = p.first + li;
   const int o = __ldg(t.orig + i);
   if (o < 0) return;
   const int ih = REDUCED ? __ldg(t.ham + i) : 0;
   SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
   SpinVec out;
   if (mc_update_site<REDUCED>(t, p, S, i, k, o, ih, smc, smd, smb, out)) {
      S[i] = out;
      if (accepted) atomicAdd(accepted, 1u);
   }
}

// Small colour classes / long neighbour lists (e.g. FeCo B2: z = 258, 114 colours): LPA lanes share one update.  They
// split the neighbour list of the atom (atom-major table nlrow: coalesced), reduce the Heisenberg sum with warp shuffles
// (fixed butterfly: deterministic), and the first lane of the group finishes the update (DM / BQ / anisotropy terms,
// draws, acceptance).  A launch then has LPA times more threads in flight and LPA times fewer dependent gathers per
// thread: the colour launches of such systems are latency-bound, not bandwidth-bound.
// one cooperative update: the LPA lanes of a group (q = lane within the group) work on attempt (li, k) of the class
// REDUCED = false: one coupling row per atom (do_reduced N, random alloys), read from the atom-major copy t.cprow next to t.nlrow;
// rows are zero-padded beyond the atom's list length and the padding entries of nlrow point at the atom itself.
template <int LPA, bool REDUCED>
__device__ __forceinline__ void mc_coop_update(const Tables& t, const McParams& p, SpinVec* __restrict__ cur, int li, int k, int q,
                                               const double* smc, const double* smd, const double* smb) {
   SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
   int i = 0, o = -1, ih = 0, n = 0;
   if (li < p.count) {
      i = p.first + li;
      o = __ldg(t.orig + i);
      if (o >= 0) {
         if (REDUCED) { ih = __ldg(t.ham + i); n = __ldg(t.lsize + ih); }
         else n = t.z;
      }
   }
   const int* __restrict__ row = t.nlrow + (size_t)i * t.z;
   const double* __restrict__ crow = REDUCED ? (smc ? smc + (size_t)ih * t.z : t.cp + (size_t)ih * t.z) : t.cprow + (size_t)i * t.z;
   double f[3] = {0.0, 0.0, 0.0};
   for (int j0 = q; j0 < n; j0 += 4 * LPA) {
      int nb[4];
      SpinVec v[4];
      double c[4];
#pragma unroll
      for (int a = 0; a < 4; a++) nb[a] = (j0 + a * LPA < n) ? __ldg(row + j0 + a * LPA) : i;
#pragma unroll
      for (int a = 0; a < 4; a++) c[a] = (j0 + a * LPA < n) ? (REDUCED ? crow[j0 + a * LPA] : __ldg(crow + j0 + a * LPA)) : 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) v[a] = S[nb[a]];
#pragma unroll
      for (int a = 0; a < 4; a++)
         if (j0 + a * LPA < n) {
            f[0] = fma(c[a], v[a].x * v[a].m, f[0]);
            f[1] = fma(c[a], v[a].y * v[a].m, f[1]);
            f[2] = fma(c[a], v[a].z * v[a].m, f[2]);
         }
   }
#pragma unroll
   for (int off = LPA / 2; off > 0; off >>= 1) {
      f[0] += __shfl_xor_sync(0xffffffffu, f[0], off);
      f[1] += __shfl_xor_sync(0xffffffffu, f[1], off);
      f[2] += __shfl_xor_sync(0xffffffffu, f[2], off);
   }
   if (q == 0 && o >= 0) {
      SpinVec out;
      if (mc_update_site<REDUCED, false>(t, p, S, i, k, o, ih, smc, smd, smb, out, f)) S[i] = out;
   }
}
template <int LPA, bool REDUCED>
__global__ void __launch_bounds__(256)
mc_colour_coop_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, SpinVec* __restrict__ cur) {
   extern __shared__ double sm[];
   const double *smc = nullptr, *smd = nullptr, *smb = nullptr;
   if (REDUCED) stage_couplings(t, sm, smc, smd, smb);
   const int gt = blockIdx.x * blockDim.x + threadIdx.x;
   mc_coop_update<LPA, REDUCED>(t, p, cur, gt / LPA, blockIdx.y, gt % LPA, smc, smd, smb);
}
template <int LPA, bool REDUCED>
__global__ void __launch_bounds__(256)
mc_sweeps_persistent_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p0, const int2* __restrict__ classes,
                            int ncol, int nsweeps, SpinVec* __restrict__ cur) {
   namespace cg = cooperative_groups;
   cg::grid_group grid = cg::this_grid();
   extern __shared__ double sm[];
   const double *smc = nullptr, *smd = nullptr, *smb = nullptr;
   if (REDUCED) stage_couplings(t, sm, smc, smd, smb);
   McParams p = p0;
   const int q = threadIdx.x % LPA;
   const long g0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) / LPA, ng = (long)gridDim.x * blockDim.x / LPA;
   for (int s = 0; s < nsweeps; s++) {
      p.sweep = p0.sweep + (unsigned long long)s;
      for (int c = 0; c < ncol; c++) {
         const int2 cl = classes[c];
         p.first = cl.x; p.count = cl.y;
         const long total = (long)cl.y * t.M;
         // every lane of a group runs the same number of rounds (the shuffles need all of them)
         for (long a0 = 0; a0 < total; a0 += ng) {
            const long a = a0 + g0;
            const bool live = a < total;
            mc_coop_update<LPA, REDUCED>(t, p, cur, live ? (int)(a % cl.y) : p.count, live ? (int)(a / cl.y) : 0, q, smc, smd, smb);
         }
         grid.sync();
      }
   }
}

// SMALL systems (the reference's own Monte Carlo regression cases: a few hundred atoms, colour classes of a few dozen):
// every colour launch is latency-bound (~6-12 us each, 8 per sweep for bcc Fe).  One CTA per ensemble keeps the whole state
// of its ensemble in shared memory and runs ALL colours of ALL requested sweeps, separated by __syncthreads(); a CTA
// barrier does not touch L1, so the neighbour table stays there after the first sweep.  Same update function, same draws
// (keyed by atom, ensemble, sweep) as the colour launches: the chain is bit-identical.
template <bool REDUCED>
__global__ void __launch_bounds__(512, 1)
mc_resident_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p0, const int2* __restrict__ classes, int ncol,
                   int nsweeps, SpinVec* __restrict__ cur) {
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   stage_couplings(t, sm, smc, smd, smb);
   const int ncpl = t.sm_cp + t.sm_dm + t.sm_bq;
   SpinVec* __restrict__ sh = reinterpret_cast<SpinVec*>(sm + ((ncpl + 3) & ~3));
   const int k = blockIdx.x;
   SpinVec* __restrict__ curk = cur + (size_t)k * t.Npad;
   for (int i = threadIdx.x; i < t.Npad; i += blockDim.x) sh[i] = curk[i];
   __syncthreads();
   McParams p = p0;
   for (int s = 0; s < nsweeps; s++) {
      p.sweep = p0.sweep + (unsigned long long)s;
      for (int c = 0; c < ncol; c++) {
         const int2 cl = classes[c];
         for (int a = threadIdx.x; a < cl.y; a += blockDim.x) {
            const int i = cl.x + a;
            const int o = __ldg(t.orig + i);
            if (o < 0) continue;
            const int ih = REDUCED ? __ldg(t.ham + i) : 0;
            SpinVec out;
            if (mc_update_site<REDUCED>(t, p, sh, i, k, o, ih, smc, smd, smb, out)) sh[i] = out;
         }
         __syncthreads();
      }
   }
   for (int i = threadIdx.x; i < t.Npad; i += blockDim.x) curk[i] = sh[i];
}

// Lattice (brick) layout with a periodic colouring: `col[slot]` is the colour of the atom (255 = padding); a launch
// updates the atoms of colour p.colour of the tiles in `tr`.  This is the Monte Carlo path of a slab-decomposed
// supercell: EDGE launches cover the boundary tiles, store every changed boundary spin into the ring neighbours'
// halo slots as well and publish the exchange epoch -- one halo exchange per colour.
template <bool REDUCED, bool EDGE>
__global__ void __launch_bounds__(256, ASD_MC_MINB)
mc_tile_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, const __grid_constant__ EdgeParams ep,
               const TileRange tr, const unsigned char* __restrict__ col, SpinVec* __restrict__ cur) {
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   stage_couplings(t, sm, smc, smd, smb);
   const int tile = ((int)blockIdx.x < tr.split) ? tr.first + (int)blockIdx.x : tr.second + ((int)blockIdx.x - tr.split);
   const int i = tile * 256 + threadIdx.x;
   const int k = blockIdx.y;
   if (i < t.Nown && __ldg(col + i) == (unsigned char)p.colour) {
      const int o = __ldg(t.orig + i);
      const int ih = REDUCED ? __ldg(t.ham + i) : 0;
      SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
      SpinVec out;
      if (mc_update_site<REDUCED>(t, p, S, i, k, o, ih, smc, smd, smb, out)) {
         S[i] = out;
         if (EDGE) {
            const int lo = __ldg(ep.hdst_lo + i), hi = __ldg(ep.hdst_hi + i);
            if (lo >= 0) ep.peer_lo[(size_t)k * t.Npad + lo] = out;
            if (hi >= 0) ep.peer_hi[(size_t)k * t.Npad + hi] = out;
         }
      }
   }
   if (EDGE) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
         const unsigned int total = gridDim.x * gridDim.y;
         if (atomicAdd(ep.ctr, 1u) == total - 1) {
            *ep.ctr = 0;
            __threadfence_system();
            if (ep.flag_lo) st_release_sys(ep.flag_lo, ep.epoch);
            if (ep.flag_hi) st_release_sys(ep.flag_hi, ep.epoch);
         }
      }
   }
}

// colour of every owned slot of a lattice layout from the colouring of a small periodic cell (p1 x p2 x p3 cells)
__global__ void lattice_colour_kernel(const LatticeDesc d, int p1, int p2, int p3, const unsigned char* __restrict__ cellcol,
                                      unsigned char* __restrict__ col) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= d.Nown) return;
   int i0, ix, iy, iz;
   unsigned char c = 255;
   if (lattice_unslot(d, s, i0, ix, iy, iz)) {
      const int gz = d.z0 + iz;
      c = cellcol[(((gz % p3) * p2 + iy % p2) * p1 + ix % p1) * d.NA + i0];
   }
   col[s] = c;
}

}  // namespace asd
