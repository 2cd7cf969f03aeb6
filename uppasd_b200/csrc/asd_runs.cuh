// Run-compressed neighbour table + register-blocked stage kernel (the fast path of lattice layouts).
//
// What the table means is unchanged -- entry j of atom i is nlist(j,i) of the reference
// (source/Hamiltonian/hamiltoniandatatype.f90:30-33), summed with ncoup(j,aHam(i)) as in heisenberg_field
// (source/Hamiltonian/hamiltonianactions.f90:461-464) -- but two regularities of a lattice in brick order are
// used to store and to read it (both VERIFIED on the device when the table is built; a layout that lacks them keeps
// the plain staged kernel of asd_device.cuh):
//
//  (1) a warp is one 32-cell x-run of one sublattice, and the gather list of its tile is sorted by (sublattice,
//      atom index): the 32 lanes' j-th neighbours are 32 CONSECUTIVE positions of the list.  One 16-bit base per
//      (run, j) replaces 32 per-atom indices: the index stream (4z bytes per atom and stage in the reference's
//      layout, 2z in nl16) disappears from HBM traffic.
//  (2) R runs that are adjacent in y / z share most of their neighbours: for the 4-shell bcc table the 4 x 50
//      neighbour runs of a 2 x 2 block of x-runs are only 100 distinct runs.  A thread therefore owns R atoms (one
//      per run of its group), reads every distinct neighbour ONCE from shared memory and feeds it to the 1..R
//      accumulators that use it.  Shared-memory traffic -- the bound of the one-atom-per-thread kernel (ncu: L1 data
//      pipe 86 %, profiles/) -- halves; FP64 FMA count is unchanged.
//
// Per group of R runs the build kernel emits the union as entries {position base, mask of runs that use it, the
// neighbour number j of each such run}, sorted by (mask, j of the lowest run) -- a key that does not depend on how
// the supercell is decomposed, so slabs reproduce the undecomposed summation order bit for bit.  The stage kernel
// walks the entries mask by mask (15 compile-time variants of the inner loop for R = 4): no predicates, no wasted
// FMAs.  The order of summation inside one atom differs from j = 1..z (parity bar: 1e-12 relative, tests).
#pragma once
#include "asd_device.cuh"

namespace asd {

#ifndef ASD_RUN_UNROLL
#define ASD_RUN_UNROLL 4
#endif

#ifndef ASD_INT_UNROLL
#define ASD_INT_UNROLL 1
#endif
#ifndef ASD_ABL
#define ASD_ABL 0      // development only: ablation bits (1: no integrator math, 2: no union walk, 4: no staging) -- results are WRONG
#endif
#ifndef ASD_WALK_U
#define ASD_WALK_U 2
#endif
// MM layouts: every mask class of a union row is padded with null entries (zero moment) to a multiple of WALK_U entries, and the
// walk runs WALK_U entries per iteration with no remainder loop (1: no padding, the loop form of the other layouts)
constexpr int WALK_U = ASD_WALK_U;
#ifndef ASD_WALK_X
#define ASD_WALK_X 1
#endif
constexpr int INT_UNROLL = ASD_INT_UNROLL;   // atoms of a thread whose integrators are interleaved (1: one rolled loop)
constexpr int RUN_UNROLL = ASD_RUN_UNROLL;   // entries of the union walk in flight per mask loop
constexpr int RUN_MAXPAIR = 256;   // R * z must stay below this (j and the entry counts are bytes)

// One warp per group of R runs.  pass 0: gcount[g] = number of union entries, or -1 if the group is not regular;
// pass 1: writes the row  utab[g][rowlen] (16-byte words):  word 0 = end[m], m = 0..15, as bytes (one past the last entry
// whose mask is <= m), then entries {x = pos_scale * base (byte offset of the record in the staged list: 24 for the list of
// 24-byte records, 8 for the component planes of the MM kernels), y / z = byte offsets
// of the couplings of runs 0..3 in Tables::cpl_small, 16 bits each}; at least one spare (zero) entry ends the row.
template <int R>
__global__ void __launch_bounds__(32)
run_union_kernel(int Nown, int Npad, int z, const uint4* __restrict__ nl16, const int2* __restrict__ meta, const int* __restrict__ lsize,
                 int pass, int rowlen, int* __restrict__ gcount, uint4* __restrict__ utab, unsigned pos_scale, int pad, unsigned null_pos) {
   // pad > 1: every mask class is filled up to a multiple of `pad` entries with null entries {pos_scale * null_pos, 0, 0} (the
   // kernel keeps a zero moment there); pass 0 then reports (padded entries << 10) | entries
   __shared__ int pstart[17], pcnt[16];
   __shared__ unsigned short base[RUN_MAXPAIR], rj[RUN_MAXPAIR];
   __shared__ unsigned int lkey[RUN_MAXPAIR];
   __shared__ uint4 lent[RUN_MAXPAIR];
   __shared__ int nlead, dup;
   const int g = blockIdx.x, lane = threadIdx.x;
   int npair = 0, ham0 = -1;
   int why = 0;   // first reason a group is not regular (negative code reported in gcount)
   bool ok = true;
   for (int r = 0; r < R; r++) {
      const int s = (g * R + r) * 32 + lane;
      int2 mt = make_int2(-1, -1);
      if (s < Nown) mt = meta[s];
      const bool act = mt.y >= 0;
      const unsigned am = __ballot_sync(0xffffffffu, act);
      if (am == 0) continue;                              // a run of padding slots
      if (am & (am + 1)) { ok = false; if (!why) why = -1; continue; }   // the real cells must be lanes 0..n-1
      const int hr = __shfl_sync(0xffffffffu, mt.x, 0);
      if (ham0 < 0) ham0 = hr; else if (hr != ham0) { ok = false; if (!why) why = -2; }
      if (__any_sync(0xffffffffu, act && mt.x != hr)) { ok = false; if (!why) why = -2; }
      const int n = lsize[hr];
      if (npair + n >= RUN_MAXPAIR) { ok = false; if (!why) why = -3; continue; }
      for (int q = 0; 8 * q < n; q++) {
         const uint4 w = nl16[(size_t)q * Npad + s];
         const unsigned li[8] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16, w.z & 0xffffu, w.z >> 16, w.w & 0xffffu, w.w >> 16};
#pragma unroll
         for (int u = 0; u < 8; u++) {
            const int j = 8 * q + u;
            if (j < n) {
               const unsigned b = __shfl_sync(0xffffffffu, li[u], 0);
               if (__any_sync(0xffffffffu, act && li[u] != b + lane)) { ok = false; if (!why) why = -4; }
               if (lane == 0) { base[npair] = (unsigned short)b; rj[npair] = (unsigned short)((r << 8) | j); }
               npair++;
            }
         }
      }
   }
   if (lane == 0) { nlead = 0; dup = 0; }
   __syncwarp();
   for (int p0 = 0; p0 < npair; p0 += 32) {
      const int p = p0 + lane;
      if (p < npair) {
         bool leader = true;
         for (int q = 0; q < p; q++) if (base[q] == base[p]) { leader = false; break; }
         if (leader) {
            unsigned mask = 0, off[4] = {0u, 0u, 0u, 0u};
            for (int q = p; q < npair; q++)
               if (base[q] == base[p]) {
                  const unsigned r = rj[q] >> 8, j = rj[q] & 255u;
                  if (mask & (1u << r)) dup = 1;      // one atom lists the same neighbour twice: not handled here
                  mask |= 1u << r;
                  off[r] = (unsigned)(ham0 * z + (int)j) * 8u;   // byte offset of ncoup(j, aHam) in Tables::cpl_small
               }
            const int at = atomicAdd(&nlead, 1);
            lkey[at] = (mask << 8) | (rj[p] & 255u);   // pair p is the first use: lowest run, its j
            lent[at] = make_uint4((unsigned)base[p] * pos_scale, off[0] | (off[1] << 16), off[2] | (off[3] << 16), 0u);
         }
      }
   }
   __syncwarp();
   const int nl = nlead;
   if (dup) { ok = false; if (!why) why = -5; }
   if (lane < 16) {
      int c = 0;
      for (int q = 0; q < nl; q++) c += (int)(lkey[q] >> 8) == lane;
      pcnt[lane] = c;
   }
   __syncwarp();
   if (lane == 0) {
      pstart[0] = 0;
      for (int m = 0; m < 16; m++) pstart[m + 1] = pstart[m] + (pcnt[m] + pad - 1) / pad * pad;
   }
   __syncwarp();
   if (pass == 0) {
      if (lane == 0) gcount[g] = ok ? ((pstart[16] << 10) | nl) : why;
      return;
   }
   if (!ok) return;
   uint4* __restrict__ row = utab + (size_t)g * rowlen;
   for (int l = lane; l < nl; l += 32) {
      const unsigned ml = lkey[l] >> 8;
      int rank = 0;
      for (int q = 0; q < nl; q++) rank += (lkey[q] >> 8) == ml && lkey[q] < lkey[l];
      row[1 + pstart[ml] + rank] = lent[l];
   }
   if (lane < 16) {
      for (int q = pstart[lane] + pcnt[lane]; q < pstart[lane + 1]; q++) row[1 + q] = make_uint4(null_pos * pos_scale, 0u, 0u, 0u);
      reinterpret_cast<unsigned char*>(row)[lane] = (unsigned char)pstart[lane + 1];
   }
}

// inner loops of the union walk, one per mask value (ascending, like the entries)
// PL: distance in doubles between the components of one staged moment (1: 24-byte records, MM_PLANE: component planes)
constexpr int MM_PLANE = 3232;     // positions per component plane of the MM kernels (layouts with ucap + 32 <= MM_PLANE)
template <int R, int MASK, int PL = 1, int U = 1>
struct RunLoop {
   static __device__ __forceinline__ void use(const Tables& t, const uint4& en, const double* __restrict__ srec, double (&f)[R][3]) {
      const double* __restrict__ m = reinterpret_cast<const double*>(reinterpret_cast<const char*>(srec) + en.x);
      const double mx = m[0], my = m[PL], mz = m[2 * PL];
      const char* __restrict__ cb = reinterpret_cast<const char*>(t.cpl_small);
#pragma unroll
      for (int r = 0; r < R; r++)
         if (MASK & (1 << r)) {
            const unsigned w = (r < 2) ? en.y : en.z;
            const unsigned off = (r & 1) ? (w >> 16) : (w & 0xffffu);
            const double c = *reinterpret_cast<const double*>(cb + off);
            f[r][0] = fma(c, mx, f[r][0]);
            f[r][1] = fma(c, my, f[r][1]);
            f[r][2] = fma(c, mz, f[r][2]);
         }
   }
   static __device__ __forceinline__ void run(const Tables& t, const uint4* __restrict__ ent, const unsigned char* __restrict__ endb,
                                              const double* __restrict__ srec, double (&f)[R][3]) {
      RunLoop<R, MASK - 1, PL, U>::run(t, ent, endb, srec, f);
      const int e0 = endb[MASK - 1], e1 = endb[MASK];
      if (U == 1) {
         uint4 nx = ent[e0];                   // one entry ahead: its decode overlaps the loads of the current one
#pragma unroll (RUN_UNROLL)
         for (int e = e0; e < e1; e++) {
            const uint4 en = nx;
            nx = ent[e + 1];                   // rows carry one spare entry
            use(t, en, srec, f);
         }
      } else if (U == 2 && ASD_WALK_X) {
         // classes padded to an even number of entries (null entries read a zero moment): four entries per iteration -- the first
         // pair was fetched by the previous iteration, the second pair flies while the first is used -- and one closing pair
         int e = e0;
         uint4 n0 = ent[e], n1 = ent[e + 1];
#pragma unroll 1
         for (; e + 4 <= e1; e += 4) {
            const uint4 b0 = ent[e + 2], b1 = ent[e + 3];
            use(t, n0, srec, f); use(t, n1, srec, f);
            n0 = ent[e + 4]; n1 = ent[e + 5];        // rows carry two spare entries
            use(t, b0, srec, f); use(t, b1, srec, f);
         }
         if (e < e1) { use(t, n0, srec, f); use(t, n1, srec, f); }
      } else {
         // classes padded to a multiple of U entries (null entries read a zero moment): U entries per iteration, the next U
         // fetched one iteration ahead (rows carry U spare entries), no remainder loop
         uint4 nx[U];
#pragma unroll
         for (int u = 0; u < U; u++) nx[u] = ent[e0 + u];
#pragma unroll 1
         for (int e = e0; e < e1; e += U) {
            uint4 en[U];
#pragma unroll
            for (int u = 0; u < U; u++) { en[u] = nx[u]; nx[u] = ent[e + U + u]; }
#pragma unroll
            for (int u = 0; u < U; u++) use(t, en[u], srec, f);
         }
      }
   }
};
template <int R, int PL, int U>
struct RunLoop<R, 0, PL, U> {
   static __device__ __forceinline__ void run(const Tables&, const uint4*, const unsigned char*, const double*, double (&)[R][3]) {}
};

// One stage of one LLG step on a tile of NW*128 slots = NW groups of R = 4 x-runs (NW = 2, 4, 8: tiles of 256, 512,
// 1024 slots = 1, 2, 4 bricks of a super-brick): NW warps, thread (warp w, lane l) owns the atoms
// tile*TS + (w*4 + r)*32 + l, r = 0..3.  Same contract as llg_stage_kernel (asd_device.cuh).
// XS: the layout's DM / BQ neighbours are in the gather list too (dm16 / bq16) and are read from shared memory.
// LEAN (1; 2 = with single-ion anisotropy; 3 = general field terms, lean integrator only): Heisenberg system with a uniform field and uniform LLG parameters (the engine checks: no DM / BQ /
// tensor couplings, uniform external field, uniform damping / g factor / temperature, no torque field, mompar 0) -- the integrator
// loop then carries none of the runtime checks and predicated loads of the general form (270 -> 150 instructions per atom-stage).
// MM: the gather list is staged from the MOMENT PLANES -- emomM = e * m of every slot as three component planes [M][3][Npad]
// (p.mm_cur / p.mm_pred), which every MM launch keeps up to date next to the spins it writes -- with 8-byte asynchronous copies
// (cp.async) straight into component planes in shared memory: no registers, no conversion, all copies of a thread in flight at
// once, 24 instead of 32 bytes per gathered spin.  (Ablation, profiles/README: staging through registers cost 0.10 of the 0.47 ms
// step and did not overlap with anything.)  Layouts with ucap + 64 <= MM_PLANE, no XS tables (a slab pushes the planes of its boundary atoms too); the union rows carry
// 8 * base.
template <int SOLVER, int STAGE, int NW, bool EDGE, bool MSUM, bool XS, int LEAN = 0, bool MM = false>
__global__ void __launch_bounds__(NW * 32, (NW == 8) ? 2 : (NW == 4) ? 4 : 6)
llg_runs_kernel(const __grid_constant__ Tables t, const __grid_constant__ LlgParams p, const __grid_constant__ EdgeParams ep,
                const TileRange tr, SpinVec* __restrict__ cur, SpinVec* __restrict__ pred, double* __restrict__ b2eff) {
   constexpr int R = 4, NT = NW * 32, TS = NW * R * 32;
   constexpr int SB = (NW == 8) ? 7 : 8;   // spins in flight per thread while the gather list is staged
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   const int tile = ((int)blockIdx.x < tr.split) ? tr.first + (int)blockIdx.x : tr.second + ((int)blockIdx.x - tr.split);
   const int k = blockIdx.y;
   const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
   SpinVec* __restrict__ curk = cur + (size_t)k * t.Npad;
   SpinVec* __restrict__ predk = pred + (size_t)k * t.Npad;
   const SpinVec* __restrict__ S = (STAGE == 1) ? curk : predk;
   const int i0 = tile * TS + wp * (R * 32) + ln;      // atom of run r: i0 + 32 r
   // ---- every independent global load first: {ham, orig}, own spins, union rows, first batch of the gather list ----
   int2 mt[R];
#pragma unroll
   for (int r = 0; r < R; r++) {
      const int i = i0 + 32 * r;
      const int ii = (i < t.Nown) ? i : 0;
      mt[r] = __ldg(t.meta + ii);
      if (i >= t.Nown) mt[r].y = -1;
   }
   // XS: the first DM position word of the 4 atoms (one 16-byte load each) joins the independent loads of the prologue; fetched
   // inside the integrator loop it was a dependent global load per atom (ncu r2p: 10 % of the stall samples of config 4)
   uint4 dmw[R];
   if (XS) {
#pragma unroll
      for (int r = 0; r < R; r++) {
         const int i = i0 + 32 * r;
         dmw[r] = (t.dm16 != nullptr && i < t.Nown) ? __ldg(t.dm16 + i) : make_uint4(0u, 0u, 0u, 0u);
      }
   }
   const int cnt = __ldg(t.ucount + tile);
   const int* __restrict__ ul = t.ulist + (size_t)tile * t.ucap;
   const int ncpl = t.sm_dm + t.sm_bq;                       // exchange couplings ride in the constant bank
   double* __restrict__ s3 = sm + ((ncpl + 1) & ~1);          // keeps the union rows 16-byte aligned
   uint4* __restrict__ rows = reinterpret_cast<uint4*>(s3 + 3 * (MM ? MM_PLANE : (t.ucap + 32)));
   const int nrow = NW * t.urow;
   const uint4* __restrict__ rsrc = t.utab + (size_t)tile * nrow;
   // union rows: asynchronous 16-byte copies straight into shared memory (no registers held while they fly)
   for (int q = threadIdx.x; q < nrow; q += NT) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(rows + q);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rsrc + q) : "memory");
   }
   asm volatile("cp.async.commit_group;" ::: "memory");
   // Programmatic dependent launch: everything above reads tables only.  Let the next stage's CTAs be scheduled as soon as
   // every CTA of this grid has started, and do not touch a spin before the previous stage has completed and flushed.
   // (Both are no-ops for an ordinary launch.)
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
   asm volatile("griddepcontrol.wait;" ::: "memory");
   // L2 prefetch of what the CTA one wave later needs first: its own spins, gather list and union rows
   if (t.pf_tiles > 0 && threadIdx.x < (MM ? 6 : 3)) {
      const size_t nt = (size_t)tile + t.pf_tiles;
      if (nt * TS < (size_t)t.Nown) {
         if (MM && threadIdx.x >= 3) {
            // the tile's own part of the moment planes (a third of its gather list)
            const double* __restrict__ G = ((STAGE == 1) ? p.mm_cur : p.mm_pred) + ((size_t)k * 3 + (threadIdx.x - 3)) * t.Npad + nt * TS;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G), "r"((unsigned)(TS * 8)) : "memory");
         }
         if (threadIdx.x == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(S + nt * TS), "r"((unsigned)(TS * 32)) : "memory");
         if (threadIdx.x == 1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(t.ulist + nt * t.ucap), "r"((unsigned)t.ucap * 4u) : "memory");
         if (threadIdx.x == 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(t.utab + nt * nrow), "r"((unsigned)nrow * 16u) : "memory");
      }
   }
   float gn[R][3];
#pragma unroll
   for (int r = 0; r < R; r++) { gn[r][0] = 0.f; gn[r][1] = 0.f; gn[r][2] = 0.f; }
   // short gather lists (layouts with few neighbours: DM systems, 2-D lattices): rounds of 2 x 3 spins per thread -- the long form
   // below executes its 14 predicated load / convert / store slots whatever the list length (ncu r2s: 670 of 3770 instructions of
   // config 4).  Only in the XS kernels, so that the code of the headline kernel is not touched.
   if (MM) {
      // the zero moment the null entries of padded union rows read: positions MM_PLANE - 32 .. MM_PLANE - 1 of every plane
      if (WALK_U > 1 && threadIdx.x < 96) s3[(threadIdx.x >> 5) * MM_PLANE + (MM_PLANE - 32) + (threadIdx.x & 31)] = 0.0;
      const double* __restrict__ G = ((STAGE == 1) ? p.mm_cur : p.mm_pred) + (size_t)k * 3 * t.Npad;
      constexpr int MB = 14;    // list positions per thread and round: every index first, then 3 copies per position
      for (int u0 = threadIdx.x; u0 < cnt; u0 += MB * NT) {
         int sl[MB];
#pragma unroll
         for (int a = 0; a < MB; a++) sl[a] = (u0 + a * NT < cnt) ? __ldg(ul + u0 + a * NT) : -1;
#pragma unroll
         for (int a = 0; a < MB; a++)
            if (sl[a] >= 0) {
               const unsigned dst = (unsigned)__cvta_generic_to_shared(s3 + u0 + a * NT);
               const double* __restrict__ g = G + sl[a];
               asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g) : "memory");
               asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * MM_PLANE), "l"(g + t.Npad) : "memory");
               asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 16u * MM_PLANE), "l"(g + 2 * (size_t)t.Npad) : "memory");
            }
      }
      if (p.thermal) {
#pragma unroll
         for (int r = 0; r < R; r++)
            gauss3f_raw(p.seed, (uint32_t)mt[r].y + t.atom_offset, (uint32_t)k + t.ens_offset, p.step, 0u, gn[r][0], gn[r][1], gn[r][2]);
      }
   } else if (XS && t.ucap <= 6 * NT) {
      constexpr int SS = 3;
      for (int u0 = threadIdx.x; u0 < cnt; u0 += 2 * SS * NT) {
         int sl[2 * SS];
#pragma unroll
         for (int a = 0; a < 2 * SS; a++) sl[a] = (u0 + a * NT < cnt) ? __ldg(ul + u0 + a * NT) : 0;
         SpinVec v[2 * SS];
#pragma unroll
         for (int a = 0; a < 2 * SS; a++) v[a] = S[sl[a]];
         if (u0 == (int)threadIdx.x && p.thermal) {
#pragma unroll
            for (int r = 0; r < R; r++)
               gauss3f_raw(p.seed, (uint32_t)mt[r].y + t.atom_offset, (uint32_t)k + t.ens_offset, p.step, 0u, gn[r][0], gn[r][1], gn[r][2]);
         }
#pragma unroll
         for (int a = 0; a < 2 * SS; a++) {
            const int u = u0 + a * NT;
            if (u < cnt) {
               double* __restrict__ m = s3 + 3 * u;
               m[0] = v[a].x * v[a].m; m[1] = v[a].y * v[a].m; m[2] = v[a].z * v[a].m;
            }
         }
      }
   } else
   // gather list -> shared memory in rounds of 2 x SB spins per thread: the indices of both halves are fetched first
   // (one round trip), the spins of the second half fly while the first half is converted and stored
   for (int u0 = threadIdx.x; u0 < ((ASD_ABL & 4) ? 0 : cnt); u0 += 2 * SB * NT) {
      int sl[2 * SB];
#pragma unroll
      for (int a = 0; a < 2 * SB; a++) sl[a] = (u0 + a * NT < cnt) ? __ldg(ul + u0 + a * NT) : 0;
      SpinVec v[SB];
#pragma unroll
      for (int a = 0; a < SB; a++) v[a] = S[sl[a]];
      if (u0 == (int)threadIdx.x && p.thermal) {
         // while the first batch is in flight: the Langevin noise of the 4 atoms (pure ALU work, independent of the field)
#pragma unroll
         for (int r = 0; r < R; r++)
            gauss3f_raw(p.seed, (uint32_t)mt[r].y + t.atom_offset, (uint32_t)k + t.ens_offset, p.step, 0u, gn[r][0], gn[r][1], gn[r][2]);
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
#pragma unroll
         for (int a = 0; a < SB; a++) {
            const int u = u0 + (h * SB + a) * NT;
            if (u < cnt) {
               double* __restrict__ m = s3 + 3 * u;
               m[0] = v[a].x * v[a].m; m[1] = v[a].y * v[a].m; m[2] = v[a].z * v[a].m;
            }
            if (h == 0) v[a] = S[sl[SB + a]];
         }
      }
   }
   asm volatile("cp.async.wait_all;" ::: "memory");
   stage_couplings(t, sm, smc, smd, smb);   // ends with __syncthreads() when it stages anything
   if (ncpl == 0) __syncthreads();
#ifndef ASD_NO_OWNPF
   if (MM) {
      // the own (and old) spin of the first atom is needed right after the union walk: bring its line into L1 now (no register is
      // held across the walk; the loads of atoms 2..4 already fly one integrator iteration ahead)
      const int ifirst = min(i0, t.Nown - 1);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(S + ifirst));
      if (STAGE == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(curk + ifirst));
   }
#endif
   // ---- Heisenberg sums of the 4 atoms of this thread: every distinct neighbour run read once ----
   double f[R][3];
#pragma unroll
   for (int r = 0; r < R; r++) { f[r][0] = 0.0; f[r][1] = 0.0; f[r][2] = 0.0; }
   int ih = -1;
#pragma unroll
   for (int r = 0; r < R; r++) ih = max(ih, __shfl_sync(0xffffffffu, mt[r].y >= 0 ? mt[r].x : -1, 0));
   if (ih >= 0 && !(ASD_ABL & 2)) {
      const uint4* __restrict__ row = rows + wp * t.urow;
      if (MM) RunLoop<R, (1 << R) - 1, MM_PLANE, WALK_U>::run(t, row + 1, reinterpret_cast<const unsigned char*>(row), s3 + ln, f);
      else RunLoop<R, (1 << R) - 1>::run(t, row + 1, reinterpret_cast<const unsigned char*>(row), s3 + 3 * ln, f);
   }
   // ---- integrators: one rolled loop over the 4 atoms (the register arrays rotate, so the body exists once); the
   //      own spin (and the old spin of a corrector) of the next atom is loaded one iteration ahead ----
   double mnew[3] = {0.0, 0.0, 0.0};
   const int ilast = t.Nown - 1;
   SpinVec own = S[min(i0, ilast)], old;
   if (STAGE == 2) old = curk[min(i0, ilast)];
#pragma unroll (INT_UNROLL)
   for (int r = 0; r < R; r++) {
      const int i = i0 + 32 * r;
      const int io = mt[0].y;
      const int inext = min(i + 32, ilast);
      const SpinVec own_n = S[inext];
      SpinVec old_n;
      if (STAGE == 2) old_n = curk[inext];
      if (io >= 0) {
         double bs[3] = {f[0][0], f[0][1], f[0][2]}, bq[3] = {0.0, 0.0, 0.0};
         double h[3];
         if (LEAN == 1 || LEAN == 2) {
            h[0] = t.hext[0]; h[1] = t.hext[1]; h[2] = t.hext[2];
            // LEAN == 2: single-ion anisotropy is the one extra term (its own instantiation: a runtime test here cost the plain
            // system 4.7 %, measured)
            if (LEAN == 2) aniso_field<true>(t, i, ih, own.x * own.m, own.y * own.m, own.z * own.m, bs[0], bs[1], bs[2], bq[0], bq[1], bq[2]);
         } else {
            site_field<true, false, ASD_CHUNK, XS>(t, S, i, ih, own, smc, smd, smb, bs, bq, s3, (XS && t.dm16 != nullptr) ? &dmw[0] : nullptr);
            ext_field(t, i, k, h);
         }
#ifndef ASD_NO_TFIELD
         h[0] += p.tf[0]; h[1] += p.tf[1]; h[2] += p.tf[2];
#endif
         double b[3];
         if (LEAN == 1) { b[0] = bs[0] + h[0]; b[1] = bs[1] + h[1]; b[2] = bs[2] + h[2]; }
         else { b[0] = bs[0] + (bq[0] + h[0]); b[1] = bs[1] + (bq[1] + h[1]); b[2] = bs[2] + (bq[2] + h[2]); }
#if (ASD_ABL & 1)
         SpinVec o = (STAGE == 1) ? own : old;
         o.x += 1e-300 * (b[0] + gn[0][0]); o.y += 1e-300 * (b[1] + gn[0][1]); o.z += 1e-300 * (b[2] + gn[0][2]);
#else
         const SpinVec o = integrate_site<SOLVER, STAGE, false, (LEAN != 0)>(t, p, i, k, io, b, own, (STAGE == 1) ? own : old, b2eff, gn[0]);
#endif
         if (STAGE == 1) predk[i] = o; else curk[i] = o;
         if (MM) {
            double* __restrict__ W = ((STAGE == 1) ? p.mm_pred : p.mm_cur) + (size_t)k * 3 * t.Npad + i;
            W[0] = o.x * o.m; W[t.Npad] = o.y * o.m; W[2 * (size_t)t.Npad] = o.z * o.m;
         }
         if (MSUM) { mnew[0] += o.x * o.m; mnew[1] += o.y * o.m; mnew[2] += o.z * o.m; }
         if (EDGE) {
            const int lo = __ldg(ep.hdst_lo + i), hi = __ldg(ep.hdst_hi + i);
            if (!MM) {
               if (lo >= 0) ep.peer_lo[(size_t)k * t.Npad + lo] = o;
               if (hi >= 0) ep.peer_hi[(size_t)k * t.Npad + hi] = o;
            } else {
               // MM layouts: the neighbours' launches gather from the moment planes only, so emomM (24 bytes) is what crosses
               // NVLink per stage; the spins of the halo are refreshed once, when asd_sd_steps returns (slab_push_state)
               const double ox = o.x * o.m, oy = o.y * o.m, oz = o.z * o.m;
               if (lo >= 0) { double* __restrict__ q = ep.peer_mlo + (size_t)k * 3 * t.Npad + lo; q[0] = ox; q[t.Npad] = oy; q[2 * (size_t)t.Npad] = oz; }
               if (hi >= 0) { double* __restrict__ q = ep.peer_mhi + (size_t)k * 3 * t.Npad + hi; q[0] = ox; q[t.Npad] = oy; q[2 * (size_t)t.Npad] = oz; }
            }
         }
      }
      own = own_n;
      if (STAGE == 2) old = old_n;
#pragma unroll
      for (int q = 0; q + 1 < R; q++) {
         f[q][0] = f[q + 1][0]; f[q][1] = f[q + 1][1]; f[q][2] = f[q + 1][2];
         mt[q] = mt[q + 1];
         gn[q][0] = gn[q + 1][0]; gn[q][1] = gn[q + 1][1]; gn[q][2] = gn[q + 1][2];
         if (XS) dmw[q] = dmw[q + 1];
      }
   }
   if (MSUM) {
      __shared__ double red[3][NW];
#pragma unroll
      for (int a = 0; a < 3; a++) {
         double v = mnew[a];
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
         if (ln == 0) red[a][wp] = v;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
         double v = 0.0;
         if (threadIdx.x < 3)
            for (int q = 0; q < NW; q++) v += red[threadIdx.x][q];
         p.msum_part[((size_t)k * p.msum_ntile + tile) * 4 + threadIdx.x] = v;
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

}  // namespace asd
