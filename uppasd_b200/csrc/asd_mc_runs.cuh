// Monte Carlo block sweep, RUN form: the block sweep of asd_mc_block.cuh for tiles whose colour classes are made of regular
// x-runs -- every interior tile of a device-built lattice with a periodic colouring.
//
// mc_block_kernel reads, per attempt, seven 16-byte words of neighbour positions, the atom's Hamiltonian row, list length,
// own spin and list position from global memory inside the colour phases; the phases are short (128 atoms of a 1024-atom tile)
// and separated by barriers, so that latency is exposed 8 times per tile (ncu r2b: long_scoreboard 5.7 of 12 cycles per issue,
// 22 instructions per neighbour).  Here the atoms of a colour class are cut into GROUPS of up to 16 consecutive entries of the
// tile's colour order.  With the x-residue split order of the gather list (asd_tiles.cuh) the 16 atoms of a group are the cells
// x0, x0 + px, x0 + 2 px ... of one line of one sublattice, and the j-th neighbours of those cells are CONSECUTIVE positions
// of the list: one 16-bit base per (group, j) replaces 16 per-atom positions, and the group shares Hamiltonian row, list
// length and a base for the atoms' own list positions.  mc_block_groups_kernel builds that table from the per-atom
// position words and VERIFIES every property per tile; a tile that fails (ragged bricks at an open or padded edge, a
// colouring whose x-period does not divide the run) keeps mc_block_kernel -- both kernels serve the tiles of one class side
// by side and run the same chain (same draws, same visiting order).
//
// The whole table of a tile (8 KB for bcc Fe, z = 50) is staged in shared memory next to the gather list, the trial moves are
// drawn for half a tile at a time into a record that also keeps |m| and the slot, and a colour phase then touches shared
// memory only: per neighbour one base extraction, 3 x LDS.64 (the moment), one constant-bank load (the coupling of the step, see
// McRuns) and 3 DFMA; two or four lanes share an atom and join with a shuffle (and, for four, a named barrier of the warp pair).
//
// Restated pieces: as asd_mc_block.cuh (choose_random_flip montecarlo_common.f90:25-79, calculate_energy :431-865,
// flip_a :190-200, flip_h :371-422).
#pragma once
#include "asd_mc_block.cuh"

namespace asd {

#ifdef ASD_MC_PROF
__device__ unsigned long long g_mc_prof[8];
#define MC_PROF_T(x) const long long x = clock64()
#define MC_PROF_ADD(i, a, b) if (threadIdx.x == 0) atomicAdd(&g_mc_prof[i], (unsigned long long)((b) - (a)))
#else
#define MC_PROF_T(x)
#define MC_PROF_ADD(i, a, b)
#endif

constexpr int MCR_GMAX = 80;      // groups per tile the table has room for (64 for a full 1024-slot tile of 16-atom groups)
constexpr int MCR_BATCH = 512;    // attempts whose draws are resident at a time (256-thread kernel)
constexpr int MCR_BATCH_WS = 256; // same, warp-specialised kernel (two record buffers)
constexpr int MCR_HDR = 8;        // 16-bit words of a group header: pos0, cnt, selfbase, ham, steps, -, -, -
constexpr int MCR_NBMAX = 6;      // draw batches per tile
constexpr int MCR_PHMAX = 62;     // colour phases per tile (colours + batch splits) the row header has room for
constexpr int MCR_PH0 = 8;        // first 16-bit word of the phase list in the row header; word 0 = number of phases
constexpr int MCR_CSEQ = 128;     // coupling-sequence entries (all Hamiltonian rows) that ride in the kernel parameters
constexpr int MCR_ZMAX = 128;     // longest neighbour list the arrangement kernel sorts

// Neighbour ARRANGEMENT of a Hamiltonian row: Q lanes share an atom, and in step s lane-slot q handles neighbour
// arr[row][q][s].  Neighbours with bit-identical couplings are dealt to the Q slots of the same step, so that the coupling of
// a step -- cseq[row][s] -- is the same for every lane of the warp (a constant-bank operand, no shared-memory read); a step
// that a coupling value cannot fill is padded with entries that point at zero records behind the gather list.  Exchange
// tables of lattices with inversion symmetry pair up exactly (J(r) = J(-r)); bcc Fe with 8 / 6 / 12 / 24 neighbours per
// shell needs 25 steps for Q = 2 and 13 for Q = 4.
struct McRuns {
   int gstride;                   // 16-bit words per tile row of gtab
   int gwords;                    // leading words of a row the sweep kernel stages (largest regular tile; multiple of 8)
   int q;                         // lanes per atom: 2 (256 threads) or 4 (512 threads, two warps per group)
   int sp;                        // steps per lane-slot, padded to a multiple of 4 (bases per slot in a group record)
   int ncw;                       // 16-bit words of the row header (multiple of 8): nph, -, ... | phases x {first group | new
                                  // batch << 15, end group, end group of the batch, -}: the colour phases in sweep order
   int zero;                      // list position of the 16 zero records ( = ucap)
   int batch;                     // attempts per draw batch the phase list is cut for (MCR_BATCH or MCR_BATCH_WS)
   const unsigned short* __restrict__ gtab;   // [ntile][gstride]: header | groups x {header, bases[q][sp]}
   const double* __restrict__ drec;           // [M][ntile][1024][5] trial-move records of THIS sweep drawn by mc_predraw_kernel (colour
                                              // order of the tile), or null: the sweep kernel draws them itself (mc_draw_batch)
   int ntile;
   double cseq[MCR_CSEQ];         // [NH][sp] coupling of every step, 0 beyond the row's last step
};

// one thread per Hamiltonian row: arrangement of its neighbour list for Q lanes per atom
__global__ void mc_runs_arrange_kernel(int NH, int z, int Q, int spcap, const int* __restrict__ lsize, const double* __restrict__ cp,
                                       unsigned short* __restrict__ arr, double* __restrict__ cseq, int* __restrict__ steps) {
   const int ih = blockIdx.x * blockDim.x + threadIdx.x;
   if (ih >= NH) return;
   const int n = min(lsize[ih], z);
   if (n > MCR_ZMAX) { steps[ih] = -1; return; }
   unsigned char idx[MCR_ZMAX];
   for (int j = 0; j < n; j++) idx[j] = (unsigned char)j;
   // insertion sort by the coupling's bit pattern (stable: equal couplings keep the list order)
   for (int a = 1; a < n; a++) {
      const unsigned char v = idx[a];
      const long long kv = __double_as_longlong(cp[(size_t)ih * z + v]);
      int b = a - 1;
      while (b >= 0 && __double_as_longlong(cp[(size_t)ih * z + idx[b]]) > kv) { idx[b + 1] = idx[b]; b--; }
      idx[b + 1] = v;
   }
   for (int q = 0; q < Q; q++)
      for (int s = 0; s < spcap; s++) arr[((size_t)ih * Q + q) * spcap + s] = 0xffffu;
   for (int s = 0; s < spcap; s++) cseq[(size_t)ih * spcap + s] = 0.0;
   int s = 0, a = 0;
   while (a < n) {
      int b = a;
      const long long kv = __double_as_longlong(cp[(size_t)ih * z + idx[a]]);
      while (b < n && __double_as_longlong(cp[(size_t)ih * z + idx[b]]) == kv) b++;
      for (int c = a; c < b; c += Q, s++) {
         if (s >= spcap) { steps[ih] = -1; return; }
         cseq[(size_t)ih * spcap + s] = cp[(size_t)ih * z + idx[a]];
         for (int q = 0; q < Q && c + q < b; q++) arr[((size_t)ih * Q + q) * spcap + s] = idx[c + q];
      }
      a = b;
   }
   steps[ih] = s;
}

// One CTA per tile: group table + verification.  ok[tile] = number of groups iff every group of the tile is regular, else 0.
__global__ void __launch_bounds__(256)
mc_block_groups_kernel(int ts, int ncol, int z, size_t Npad, const int* __restrict__ ham, const uint4* __restrict__ nl16,
                       const unsigned short* __restrict__ selfpos, const unsigned short* __restrict__ corder,
                       const int* __restrict__ cstart, const int* __restrict__ ucount, const unsigned short* __restrict__ arr,
                       const int* __restrict__ steps, int spcap, McRuns mr, unsigned short* __restrict__ gtab, int* __restrict__ ok) {
   __shared__ int cgs[66], bad;
   const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int* __restrict__ cs = cstart + (size_t)tile * (ncol + 1);
   unsigned short* __restrict__ row = gtab + (size_t)tile * mr.gstride;
   const int gw = MCR_HDR + mr.q * mr.sp;
   if (tid == 0) {
      bad = 0;
      int g = 0;
      for (int c = 0; c < ncol; c++) { cgs[c] = g; g += (cs[c + 1] - cs[c] + 15) / 16; }
      cgs[ncol] = g;
      if (g > MCR_GMAX || ncol > 64) bad = 1;
   }
   __syncthreads();
   if (bad) { if (tid == 0) ok[tile] = 0; return; }
   for (int q = tid; q < mr.ncw; q += 256) row[q] = 0;
   __syncthreads();
   const int G = cgs[ncol];
   const int cnt_list = ucount[tile];
   for (int c = 0; c < ncol; c++) {
      const int ng = cgs[c + 1] - cgs[c];
      for (int qg = warp; qg < ng; qg += 8) {
         const int pos0 = cs[c] + 16 * qg, cnt = min(16, cs[c + 1] - pos0);
         unsigned short* __restrict__ rec = row + mr.ncw + (size_t)(cgs[c] + qg) * gw;
         const bool act = lane < cnt;
         const int slot = act ? (int)corder[(size_t)tile * ts + pos0 + lane] : 0;
         const int i = tile * ts + slot;
         const int sp = act ? (int)selfpos[i] : 0, hm = act ? ham[i] : 0;
         const int sp0 = __shfl_sync(0xffffffffu, sp, 0), hm0 = __shfl_sync(0xffffffffu, hm, 0);
         bool good = !act || (slot != 0xffff && sp == sp0 + lane && hm == hm0 && hm >= 0);
         const int S = (hm0 >= 0) ? steps[hm0] : -1;
         if (S < 0 || S > mr.sp) good = false;
         if (sp0 + cnt > cnt_list) good = false;
         for (int q = 0; q < mr.q; q++)
            for (int s = 0; s < mr.sp; s++) {
               const int j = (hm0 >= 0 && s < spcap) ? (int)arr[((size_t)hm0 * mr.q + q) * spcap + s] : 0xffff;
               int b0 = mr.zero;
               if (j != 0xffff) {
                  unsigned li[8];
                  uint4 w = make_uint4(0u, 0u, 0u, 0u);
                  if (act) w = nl16[(size_t)(j >> 3) * Npad + i];
                  unpack16(w, li);
                  const int mine = (int)li[j & 7];
                  b0 = __shfl_sync(0xffffffffu, mine, 0);
                  if (act && mine != b0 + lane) good = false;
                  if (b0 + cnt > cnt_list) good = false;
               }
               if (lane == 0) rec[MCR_HDR + q * mr.sp + s] = (unsigned short)b0;
            }
         if (lane == 0) {
            rec[0] = (unsigned short)pos0; rec[1] = (unsigned short)cnt; rec[2] = (unsigned short)sp0; rec[3] = (unsigned short)hm0;
            rec[4] = (unsigned short)max(S, 0); rec[5] = 0; rec[6] = 0; rec[7] = 0;
         }
         if (!__all_sync(0xffffffffu, good)) bad = 1;
      }
   }
   // unused group records: zero (the sweep kernel stages the leading gwords of the row)
   for (int q = mr.ncw + G * gw + tid; q < mr.gstride; q += 256) row[q] = 0;
   __syncthreads();
   if (tid == 0) {
      // draw batches: consecutive groups whose attempts fit mr.batch records; phases: the colours of a batch, in order
      int nbt = 0, g = 0, nph = 0;
      while (g < G && nbt < MCR_NBMAX && !bad) {
         nbt++;
         const int gs = g;
         const int pb = row[mr.ncw + (size_t)g * gw];
         while (g < G && (int)row[mr.ncw + (size_t)g * gw] + (int)row[mr.ncw + (size_t)g * gw + 1] - pb <= mr.batch) g++;
         bool first = true;
         for (int c = 0; c < ncol; c++) {
            const int ga = max(cgs[c], gs), gb = min(cgs[c + 1], g);
            if (ga >= gb) continue;
            if (nph >= MCR_PHMAX) { bad = 1; break; }
            unsigned short* __restrict__ ph = row + MCR_PH0 + 4 * nph++;
            ph[0] = (unsigned short)(ga | (first ? 0x8000 : 0)); ph[1] = (unsigned short)gb; ph[2] = (unsigned short)g; ph[3] = 0;
            first = false;
         }
      }
      if (g < G) bad = 1;
      row[0] = (unsigned short)nph;
      ok[tile] = bad ? 0 : G;      // number of groups of a regular tile, 0: the tile keeps mc_block_kernel
   }
}

// Whole-sweep scheduling (TICKET = true): ONE launch per sweep instead of one per tile class.  Every CTA draws a ticket; tickets
// number the (tile, ensemble) pairs class by class, so a CTA only ever depends on lower tickets -- the neighbour tiles of
// LOWER classes, which must have finished this sweep before their spins are gathered (and before this tile overwrites spins
// they gather).  done[k][tile] holds the epoch of the last sweep the tile completed (st.release.gpu after a __threadfence of
// every writer; the waiters poll with ld.acquire.gpu).  A lower ticket is always held by a CTA that is already running, so the
// waits cannot deadlock whatever order the hardware starts the CTAs in.  What this buys: no partially filled last wave per
// class (bcc 128^3: 512 tiles per class on 296 CTA slots), no launch gaps, and co-resident CTAs drift out of phase, so the
// gather of one tile (HBM / L2), the trial moves of a second (FP64 pipe) and the colour phases of a third (shared-memory
// bandwidth) overlap instead of every SM doing the same phase at the same time.
struct McTicket {
   unsigned long long* __restrict__ counter;   // tickets handed out so far (all launches)
   unsigned long long base;                    // first ticket of this launch
   unsigned int epoch;                         // sweep number written to done[]
   unsigned int* __restrict__ done;            // [M][ntile]
   const int* __restrict__ adj;                // [ntile][cap] tiles whose spins the tile gathers / that gather its spins
   const int* __restrict__ nadj;               // [ntile]
   const unsigned char* __restrict__ tclass;   // [ntile] class of every tile
   int cap, ntile;
   int lookahead;                              // tickets ahead whose tile is prefetched into L2 ( = CTAs resident on the GPU)
};

__device__ __forceinline__ SpinVec ld_spin_cg(const SpinVec* p) {
   // L2-coherent load: within one launch another SM may have written the spin, and this SM's L1 may hold the old line
   SpinVec v;
   asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.m) : "l"(p) : "memory");
   return v;
}

// ---- pieces shared by the sweep kernels -------------------------------------------------------------------------------------
struct McTileCtx {
   int tile, k;
   SpinVec* __restrict__ S;                       // spins of ensemble k
   const unsigned short* __restrict__ co;         // colour order of the tile
   const unsigned short* __restrict__ grec;       // group records
   int gw, sp;
   double* __restrict__ s3;                       // emomM of the gather list
   double beta_h, beta_m;
};

// (B) draws of the attempts [pb, pe) of the colour order into rec / rslot, by ND threads (this one is number dtid); bar() is a
// barrier of exactly those threads.  Own spins: only this tile writes them in this sweep, and an atom is drawn before its visit.
//   Metropolis: choose_random_flip picks one of three trial moves per attempt; a warp whose lanes disagree would run all three
//   (sincos + sqrt | Philox + Box-Muller + sqrt + divisions | negation).  Pass 1 draws the uniforms and files every attempt
//   under its move type, pass 2 works through the three lists with all lanes on the same path.  Two attempts per thread and
//   round in pass 1: their dependent global loads (slot -> original index, own spin) fly together.
template <bool HB, int ND, class Bar>
__device__ __forceinline__ void mc_draw_batch(const Tables& t, const McParams& p, const McTileCtx& c, int dtid, int pb, int pe,
                                              double* __restrict__ rec, unsigned short* __restrict__ rslot,
                                              unsigned short* __restrict__ blist, int* __restrict__ bcnt, int bstride, Bar bar) {
   constexpr int TS = 1024, RW = 5;
   const double pi = 3.141592653589793;
   if (!HB) {
      if (dtid < 3) bcnt[dtid] = 0;
      bar();
   }
#pragma unroll 1
   for (int base = pb + dtid; base < pe; base += 2 * ND) {
      int slot[2], o[2];
      SpinVec own[2];
      bool val[2];
#pragma unroll
      for (int a = 0; a < 2; a++) {
         val[a] = base + a * ND < pe;
         slot[a] = val[a] ? (int)c.co[base + a * ND] : 0;
      }
#pragma unroll
      for (int a = 0; a < 2; a++) {
         const int i = c.tile * TS + slot[a];
         o[a] = __ldg(t.orig + i);
         own[a] = ld_spin_cg(c.S + i);
      }
#pragma unroll
      for (int a = 0; a < 2; a++) {
         if (!val[a]) continue;
         double u[4];
         uniform4(p.seed, (uint32_t)o[a] + t.atom_offset, (uint32_t)c.k + t.ens_offset, p.sweep, 1u, u);
         const int r = base + a * ND - pb;
         double* __restrict__ dr = rec + RW * r;
         if (HB) {
            double sphi, cphi;
            sincos(pi * (2.0 * u[1] - 1.0), &sphi, &cphi);
            dr[0] = u[0]; dr[1] = cphi; dr[2] = sphi; dr[3] = 0.0;
         } else {
            const int ftype = min((int)floor(3.0 * u[0]), 2);
            if (ftype == 0) { dr[0] = u[1]; dr[1] = u[2]; }
            else { dr[0] = own[a].x; dr[1] = own[a].y; dr[2] = own[a].z; }
            dr[3] = u[3];
            blist[ftype * bstride + atomicAdd(&bcnt[ftype], 1)] = (unsigned short)r;
         }
         dr[4] = own[a].m;
         rslot[r] = (unsigned short)slot[a];
      }
   }
   if (!HB) {
      bar();
      const int n0 = bcnt[0], n1 = bcnt[1], n2 = bcnt[2];
#pragma unroll 1
      for (int q = dtid; q < n0; q += ND) {
         double* __restrict__ dr = rec + RW * (int)blist[q];
         double sphi, cphi;
         sincos(dr[0] * 2 * pi, &sphi, &cphi);
         const double ct = 1.0 - 2.0 * dr[1];
         const double st = sqrt(fmax(1.0 - ct * ct, 0.0));
         dr[0] = st * cphi; dr[1] = st * sphi; dr[2] = ct;
      }
#pragma unroll 1
      for (int q = dtid; q < n1; q += ND) {
         const int r = blist[bstride + q];
         double* __restrict__ dr = rec + RW * r;
         const int o = __ldg(t.orig + c.tile * TS + (int)rslot[r]);
         double ga, gb, gc;
         gauss3f(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)c.k + t.ens_offset, p.sweep, 2u, ga, gb, gc);
         const double ax = dr[0] + ga * p.delta, ay = dr[1] + gb * p.delta, az = dr[2] + gc * p.delta;
         const double len = sqrt(ax * ax + ay * ay + az * az);
         const double rl = 1.0 / len; dr[0] = ax * rl; dr[1] = ay * rl; dr[2] = az * rl;
      }
      for (int q = dtid; q < n2; q += ND) {
         double* __restrict__ dr = rec + RW * (int)blist[2 * bstride + q];
         dr[0] = -dr[0]; dr[1] = -dr[1]; dr[2] = -dr[2];
      }
   }
}

// The trial moves of a sweep depend on the draws and on the atom's OWN spin at the start of the sweep only (an atom is drawn before its
// one visit), so they can be produced for the whole lattice by an ordinary data-parallel launch at full occupancy instead of inside the
// sweep CTAs, where 16 warps per SM cannot hide their dependent-instruction latency (21 k of the 60 k cycles of a tile, profiles/README).
// One thread per (ensemble, tile, position in the tile's colour order); same expressions as mc_draw_batch.
template <bool HB>
__global__ void __launch_bounds__(256)
mc_predraw_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, const unsigned short* __restrict__ corder, int ntile,
                  const SpinVec* __restrict__ cur, double* __restrict__ drec) {
   constexpr int TS = 1024, RW = 5;
   const double pi = 3.141592653589793;
   const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // tile * 1024 + position
   const int k = blockIdx.y;
   if (g >= (size_t)ntile * TS) return;
   const unsigned slot = corder[g];
   if (slot == 0xffffu) return;
   const int i = (int)(g / TS) * TS + (int)slot;
   const int o = __ldg(t.orig + i);
   const SpinVec own = cur[(size_t)k * t.Npad + i];
   double u[4];
   uniform4(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 1u, u);
   double d0, d1, d2, d3;
   if (HB) {
      double sphi, cphi;
      sincos(pi * (2.0 * u[1] - 1.0), &sphi, &cphi);
      d0 = u[0]; d1 = cphi; d2 = sphi; d3 = 0.0;
   } else {
      const int ftype = min((int)floor(3.0 * u[0]), 2);
      d3 = u[3];
      if (ftype == 0) {
         double sphi, cphi;
         sincos(u[1] * 2 * pi, &sphi, &cphi);
         const double ct = 1.0 - 2.0 * u[2];
         const double st = sqrt(fmax(1.0 - ct * ct, 0.0));
         d0 = st * cphi; d1 = st * sphi; d2 = ct;
      } else if (ftype == 1) {
         double ga, gb, gc;
         gauss3f(p.seed, (uint32_t)o + t.atom_offset, (uint32_t)k + t.ens_offset, p.sweep, 2u, ga, gb, gc);
         const double ax = own.x + ga * p.delta, ay = own.y + gb * p.delta, az = own.z + gc * p.delta;
         const double len = sqrt(ax * ax + ay * ay + az * az);
         const double rl = 1.0 / len; d0 = ax * rl; d1 = ay * rl; d2 = az * rl;
      } else {
         d0 = -own.x; d1 = -own.y; d2 = -own.z;
      }
   }
   double* __restrict__ dr = drec + ((size_t)k * ntile * TS + g) * RW;
   dr[0] = d0; dr[1] = d1; dr[2] = d2; dr[3] = d3; dr[4] = own.m;
}

// records of the attempts [pb, pe) of the tile's colour order from the pre-drawn array into rec / rslot (ND threads, this one dtid)
template <int ND>
__device__ __forceinline__ void mc_load_batch(const McTileCtx& c, const double* __restrict__ dtile, int dtid, int pb, int pe,
                                              double* __restrict__ rec, unsigned short* __restrict__ rslot) {
   constexpr int RW = 5;
   const int n = pe - pb;
   const double* __restrict__ src = dtile + (size_t)RW * pb;
#pragma unroll 5
   for (int q = dtid; q < RW * n; q += ND) rec[q] = __ldg(src + q);
   for (int r = dtid; r < n; r += ND) rslot[r] = c.co[pb + r];
}

// (C) one group of <= 16 same-colour atoms, handled by one warp: lanes l and l + 16 share atom l (each sums the steps of its lane-slot,
// one shuffle joins them), lanes 0..15 then decide (calculate_energy + flip_a, or flip_h) and store.  pb: first attempt of the batch
// whose records are in rec / rslot.
template <bool HB>
__device__ __forceinline__ void mc_sweep_group(const Tables& t, const McParams& p, const McRuns& mr, const McTileCtx& c, int gq, int lane,
                                               int pb, const double* __restrict__ rec, const unsigned short* __restrict__ rslot) {
   constexpr int TS = 1024, RW = 5;
   constexpr unsigned FULL = 0xffffffffu;
   const int l = lane & 15, half = lane >> 4;
   const unsigned short* __restrict__ gr = c.grec + (size_t)gq * c.gw;
   const uint2 hd = *reinterpret_cast<const uint2*>(gr);
   const int pos0 = hd.x & 0xffffu, gcnt = hd.x >> 16, sp0 = hd.y & 0xffffu, ih = hd.y >> 16, ns = gr[4];
   const bool act = l < gcnt;
   const int le = act ? l : 0;
   const double* __restrict__ lane3 = c.s3 + 3 * le;
   const unsigned short* __restrict__ bq = gr + MCR_HDR + half * c.sp;
   const int cs0 = ih * c.sp;
   double f[3] = {0.0, 0.0, 0.0};
   int s4 = 0;
#pragma unroll 2
   for (; s4 + 4 <= ns; s4 += 4) {
      const uint2 bw = *reinterpret_cast<const uint2*>(bq + s4);
      const unsigned b[4] = {bw.x & 0xffffu, bw.x >> 16, bw.y & 0xffffu, bw.y >> 16};
#pragma unroll
      for (int u = 0; u < 4; u++) {
         const double* __restrict__ m = lane3 + 3u * b[u];
         const double cc = mr.cseq[cs0 + s4 + u];
         f[0] = fma(cc, m[0], f[0]); f[1] = fma(cc, m[1], f[1]); f[2] = fma(cc, m[2], f[2]);
      }
   }
   // the last 1..3 steps one by one (bcc Fe: 25 steps = 6 blocks + 1) instead of a padded block of zero records
   for (; s4 < ns; s4++) {
      const double* __restrict__ m = lane3 + 3u * (unsigned)bq[s4];
      const double cc = mr.cseq[cs0 + s4];
      f[0] = fma(cc, m[0], f[0]); f[1] = fma(cc, m[1], f[1]); f[2] = fma(cc, m[2], f[2]);
   }
#pragma unroll
   for (int a = 0; a < 3; a++) f[a] += __shfl_xor_sync(FULL, f[a], 16);
   if (act && half == 0) {
      const int r = pos0 + l - pb;
      const double* __restrict__ dr = rec + RW * r;
      const double d0 = dr[0], d1 = dr[1], d2 = dr[2], d3 = dr[3], m = dr[4];
      double* __restrict__ mine = c.s3 + 3 * (sp0 + l);
      const double cm[3] = {mine[0], mine[1], mine[2]};
      const int i = c.tile * TS + (int)rslot[r];
      double ox, oy, oz;
      bool changed;
      if (HB) {
         // ---- flip_h: total field = beff1 + beff2 of effective_field_single; external field from the tables ----
         double bs[3] = {f[0], f[1], f[2]}, bqf[3] = {0.0, 0.0, 0.0}, h[3];
         aniso_field<true>(t, i, ih, cm[0], cm[1], cm[2], bs[0], bs[1], bs[2], bqf[0], bqf[1], bqf[2]);
         ext_field(t, i, c.k, h);
         const double tot[3] = {bs[0] + (bqf[0] + h[0]), bs[1] + (bqf[1] + h[1]), bs[2] + (bqf[2] + h[2])};
         const double zx = c.beta_h * tot[0] * p.mub * m, zy = c.beta_h * tot[1] * p.mub * m, zz = c.beta_h * tot[2] * p.mub * m;
         const double zarg = sqrt(zx * zx + zy * zy + zz * zz);
         const double rzarg = 1.0 / zarg; const double zctheta = zz * rzarg;
         const double zstheta = sqrt(1.0 - zctheta * zctheta) + 1e-14;
         const double rzs = 1.0 / (zarg * zstheta); double zcphi = zx * rzs, zsphi = zy * rzs;
         if (zx == 0.0 && zy == 0.0) { zcphi = 1.0; zsphi = 0.0; }     // degenerate frame (see mc_update_site)
         const double em2 = exp(-2.0 * zarg);
         const double ctheta = 1.0 + rzarg * log((1.0 - em2) * d0 + em2 + 1e-14);
         const double stheta = sqrt(fmax(1.0 - ctheta * ctheta, 0.0));
         const double s0 = stheta * d1, s1 = stheta * d2, s2 = ctheta;
         ox = zcphi * zctheta * s0 - zsphi * s1 + zcphi * zstheta * s2;
         oy = zsphi * zctheta * s0 + zcphi * s1 + zsphi * zstheta * s2;
         oz = -zstheta * s0 + zctheta * s2;
         changed = true;
      } else {
         // ---- calculate_energy + flip_a ----
         const double tm[3] = {d0 * m, d1 * m, d2 * m};
         double e_c = 0.0, e_t = 0.0;
         aniso_energy<true>(t, i, ih, cm, tm, e_c, e_t);
         e_c -= cm[0] * f[0] + cm[1] * f[1] + cm[2] * f[2];
         e_t -= tm[0] * f[0] + tm[1] * f[1] + tm[2] * f[2];
         e_c -= p.extfield[0] * cm[0] + p.extfield[1] * cm[1] + p.extfield[2] * cm[2];
         e_t -= p.extfield[0] * tm[0] + p.extfield[1] * tm[1] + p.extfield[2] * tm[2];
         const double de = p.mub * (e_t - e_c);
         changed = de <= 0.0 || d3 < exp(-c.beta_m * de);
         ox = d0; oy = d1; oz = d2;
      }
      if (changed) {
         SpinVec out;
         out.x = ox; out.y = oy; out.z = oz; out.m = m;
         c.S[i] = out;
         mine[0] = ox * m; mine[1] = oy * m; mine[2] = oz * m;
      }
   }
}

// gather list -> shared memory, by NT threads
template <int NT, int SB>
__device__ __forceinline__ void mc_stage_list(const McTileCtx& c, int tid, int cnt, const int* __restrict__ ul) {
   for (int u0 = tid; u0 < cnt; u0 += SB * NT) {
      int sl[SB];
#pragma unroll
      for (int a = 0; a < SB; a++) sl[a] = (u0 + a * NT < cnt) ? __ldg(ul + u0 + a * NT) : 0;
      SpinVec v[SB];
#pragma unroll
      for (int a = 0; a < SB; a++) v[a] = ld_spin_cg(c.S + sl[a]);
#pragma unroll
      for (int a = 0; a < SB; a++)
         if (u0 + a * NT < cnt) {
            double* __restrict__ m = c.s3 + 3 * (u0 + a * NT);
            m[0] = v[a].x * v[a].m; m[1] = v[a].y * v[a].m; m[2] = v[a].z * v[a].m;
         }
   }
}

// wait until every neighbour tile of a lower class has published this sweep's epoch
__device__ __forceinline__ void mc_wait_tiles(const McTicket& tk, int tile, int k, int tid) {
   const int na = __ldg(tk.nadj + tile);
   const unsigned myc = __ldg(tk.tclass + tile);
   if (tid < na) {
      const int other = __ldg(tk.adj + (size_t)tile * tk.cap + tid);
      if (__ldg(tk.tclass + other) < myc) {
         const unsigned int* f = tk.done + (size_t)k * tk.ntile + other;
         unsigned v;
         while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v == tk.epoch) break;
            __nanosleep(200);
         }
      }
   }
}

// 1024-slot tiles, 256 threads (one warp per group, two lanes per atom), reduced Hamiltonian, exchange only (no DM / BQ tables).
// HB: heat bath.  TICKET: whole-sweep scheduling (above), else one launch per tile class.
template <bool HB, bool TICKET>
__global__ void __launch_bounds__(256, 2)
mc_block_run_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, const __grid_constant__ McBlock mb,
                    const __grid_constant__ McRuns mr, const __grid_constant__ McTicket tk, const int class_first,
                    SpinVec* __restrict__ cur) {
   constexpr int NT = 256, RW = 5;      // RW: doubles per draw record {d0, d1, d2, d3, |m|}
   extern __shared__ double sm[];
   __shared__ unsigned long long s_ticket;
   __shared__ unsigned short blist[3 * MCR_BATCH];      // Metropolis: the attempts of a batch sorted by trial-move type
   __shared__ int bcnt[3];
   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   McTileCtx c;
   if (TICKET) {
      if (tid == 0) s_ticket = atomicAdd(tk.counter, 1ull) - tk.base;
      __syncthreads();
      const unsigned long long n = s_ticket;
      c.k = (int)(n % (unsigned)t.M);
      c.tile = __ldg(mb.tilelist + (int)(n / (unsigned)t.M));
   } else {
      c.tile = __ldg(mb.tilelist + class_first + (int)blockIdx.x);
      c.k = blockIdx.y;
   }
   const int tile = c.tile;
   c.S = cur + (size_t)c.k * t.Npad;
   if (TICKET && tk.lookahead > 0 && tid < 5) {
      // L2 prefetch for the CTA that will hold ticket n + lookahead (it starts about one CTA lifetime from now): what its first
      // phases read with dependent loads -- colour order, original indices and own spins (draws), group table, gather list
      const unsigned long long n2 = s_ticket + (unsigned long long)tk.lookahead;
      if (n2 < (unsigned long long)tk.ntile * (unsigned)t.M) {
         const int k2 = (int)(n2 % (unsigned)t.M);
         const size_t t2 = (size_t)__ldg(mb.tilelist + (int)(n2 / (unsigned)t.M));
         const void* ptr = nullptr;
         unsigned bytes = 0;
         if (tid == 0) { ptr = cur + (size_t)k2 * t.Npad + t2 * 1024; bytes = 1024u * 32u; }
         if (tid == 1) { ptr = mb.corder + t2 * 1024; bytes = 1024u * 2u; }
         if (tid == 2) { ptr = t.orig + t2 * 1024; bytes = 1024u * 4u; }
         if (tid == 3) { ptr = mr.gtab + t2 * mr.gstride; bytes = (unsigned)mr.gwords * 2u; }
         if (tid == 4) { ptr = mb.ulist + t2 * mb.ucap; bytes = (unsigned)mb.ucap * 4u; }
         asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes & ~15u) : "memory");
      }
   }
   MC_PROF_T(t_0);
   // shared memory: draw records [batch][5] | group table | slots [batch] | emomM of the list + 16 zero records
   double* __restrict__ rec = sm;
   unsigned short* __restrict__ gt = reinterpret_cast<unsigned short*>(rec + RW * MCR_BATCH);
   unsigned short* __restrict__ rslot = gt + mr.gwords;
   c.s3 = reinterpret_cast<double*>(rslot + MCR_BATCH);
   {
      const uint4* __restrict__ src = reinterpret_cast<const uint4*>(mr.gtab + (size_t)tile * mr.gstride);
      uint4* __restrict__ dst = reinterpret_cast<uint4*>(gt);
      for (int q = tid; q < mr.gwords / 8; q += NT) dst[q] = __ldg(src + q);
   }
   for (int q = tid; q < 48; q += NT) c.s3[3 * mr.zero + q] = 0.0;
   const int cnt = __ldg(mb.ucount + tile);
   const int* __restrict__ ul = mb.ulist + (size_t)tile * mb.ucap;
   if (TICKET) {
      // the gather list towards L2 while the first trial moves are drawn (stale lines are harmless: L2 is the coherence point)
      for (int u = tid; u < cnt; u += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(c.S + __ldg(ul + u)));
   }
   __syncthreads();
   c.co = mb.corder + (size_t)tile * 1024;
   const int nph = gt[0];
   c.grec = gt + mr.ncw;
   c.sp = mr.sp; c.gw = MCR_HDR + mr.q * mr.sp;
   c.beta_h = 1.0 / p.k_bolt / (p.temprescale * p.temperature);
   c.beta_m = 1.0 / p.k_bolt / (p.temprescale * p.temperature + 1.0e-15);
   auto bar = []() { __syncthreads(); };
   const double* __restrict__ dtile = mr.drec ? mr.drec + ((size_t)c.k * mr.ntile + tile) * (1024 * RW) : nullptr;
   auto batch_range = [&](int gs, int ge, int& pb, int& pe) {
      pb = c.grec[(size_t)gs * c.gw];
      pe = (int)c.grec[(size_t)(ge - 1) * c.gw] + (int)c.grec[(size_t)(ge - 1) * c.gw + 1];
   };
   int pb = 0, pe = 0;
   if (TICKET && nph > 0) {
      // first batch of draws BEFORE the gather: fills the wait for the neighbour tiles and the flight time of the prefetch
      batch_range(gt[MCR_PH0] & 0x7fff, gt[MCR_PH0 + 2], pb, pe);
      if (dtile) mc_load_batch<NT>(c, dtile, tid, pb, pe, rec, rslot);
      else mc_draw_batch<HB, NT>(t, p, c, tid, pb, pe, rec, rslot, blist, bcnt, MCR_BATCH, bar);
      MC_PROF_T(t_b1);
      MC_PROF_ADD(1, t_0, t_b1);
      mc_wait_tiles(tk, tile, c.k, tid);
      __syncthreads();
      MC_PROF_T(t_w);
      MC_PROF_ADD(4, t_b1, t_w);
   }
   MC_PROF_T(t_s0);
   // ---- (A) gather list -> shared memory ----
   mc_stage_list<NT, 7>(c, tid, cnt, ul);
   __syncthreads();
   MC_PROF_T(t_a);
   MC_PROF_ADD(0, t_s0, t_a);
   for (int ph = 0; ph < nph; ph++) {
      const uint2 pw = *reinterpret_cast<const uint2*>(gt + MCR_PH0 + 4 * ph);
      const int ga = pw.x & 0x7fff, gb = pw.x >> 16;
      if ((pw.x & 0x8000u) && !(TICKET && ph == 0)) {
         // ---- new batch: its draws ----
         MC_PROF_T(t_b0);
         batch_range(ga, (int)(pw.y & 0xffffu), pb, pe);
         if (dtile) mc_load_batch<NT>(c, dtile, tid, pb, pe, rec, rslot);
         else mc_draw_batch<HB, NT>(t, p, c, tid, pb, pe, rec, rslot, blist, bcnt, MCR_BATCH, bar);
         __syncthreads();
         MC_PROF_T(t_b1);
         MC_PROF_ADD(1, t_b0, t_b1);
      }
      MC_PROF_T(t_c0);
      // ---- (C) one colour of the batch; one group per warp ----
      for (int gq = ga + warp; gq < gb; gq += 8) mc_sweep_group<HB>(t, p, mr, c, gq, lane, pb, rec, rslot);
      __syncthreads();
      MC_PROF_T(t_c1);
      MC_PROF_ADD(2, t_c0, t_c1);
      if (pw.x & 0x8000u) { MC_PROF_ADD(3, t_0, t_0 + 1); }
   }
   if (TICKET) {
      // publish: every writer's stores are ordered before the flag (fence), the flag after every writer (barrier)
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(tk.done + (size_t)c.k * tk.ntile + tile), "r"(tk.epoch) : "memory");
   }
}

// WARP-SPECIALISED form (whole-sweep scheduling only): 512 threads = 8 SWEEP warps + 8 DRAW warps.  The sweep of a tile alternates
// between a phase that is bound by dependent-instruction latency (the trial moves: Philox, sincos, sqrt, divisions) and phases bound by
// shared-memory bandwidth (the neighbour loops); with two CTAs per SM (the gather list fills shared memory) the two kinds of work only
// overlap by accident.  Here the draw warps produce the records of batch b + 1 (256 attempts, double-buffered) WHILE the sweep warps
// run the colours of batch b; all 512 threads draw the first batch, wait for the neighbour tiles and stage the gather list.
// Barriers: id 1 = the 8 sweep warps (between colours), id 2 = the 8 draw warps (inside a batch), id 0 = everybody (batch switch).
template <bool HB>
__global__ void __launch_bounds__(512, 2)
mc_block_ws_kernel(const __grid_constant__ Tables t, const __grid_constant__ McParams p, const __grid_constant__ McBlock mb,
                   const __grid_constant__ McRuns mr, const __grid_constant__ McTicket tk, SpinVec* __restrict__ cur) {
   constexpr int NT = 512, RW = 5, B = MCR_BATCH_WS;
   extern __shared__ double sm[];
   __shared__ unsigned long long s_ticket;
   __shared__ unsigned short blist[3 * B];              // pass-2 lists of the batch being drawn
   __shared__ int bcnt[3];
   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const bool drawer = warp >= 8;
   McTileCtx c;
   if (tid == 0) s_ticket = atomicAdd(tk.counter, 1ull) - tk.base;
   __syncthreads();
   {
      const unsigned long long n = s_ticket;
      c.k = (int)(n % (unsigned)t.M);
      c.tile = __ldg(mb.tilelist + (int)(n / (unsigned)t.M));
   }
   const int tile = c.tile;
   c.S = cur + (size_t)c.k * t.Npad;
   // shared memory: draw records [2][B][5] | group table | slots [2][B] | emomM of the list + 16 zero records
   double* __restrict__ rec = sm;
   unsigned short* __restrict__ gt = reinterpret_cast<unsigned short*>(rec + 2 * RW * B);
   unsigned short* __restrict__ rslot = gt + mr.gwords;
   c.s3 = reinterpret_cast<double*>(rslot + 2 * B);
   {
      const uint4* __restrict__ src = reinterpret_cast<const uint4*>(mr.gtab + (size_t)tile * mr.gstride);
      uint4* __restrict__ dst = reinterpret_cast<uint4*>(gt);
      for (int q = tid; q < mr.gwords / 8; q += NT) dst[q] = __ldg(src + q);
   }
   for (int q = tid; q < 48; q += NT) c.s3[3 * mr.zero + q] = 0.0;
   const int cnt = __ldg(mb.ucount + tile);
   const int* __restrict__ ul = mb.ulist + (size_t)tile * mb.ucap;
   for (int u = tid; u < cnt; u += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(c.S + __ldg(ul + u)));
   __syncthreads();
   c.co = mb.corder + (size_t)tile * 1024;
   const int nph = gt[0];
   c.grec = gt + mr.ncw;
   c.sp = mr.sp; c.gw = MCR_HDR + mr.q * mr.sp;
   c.beta_h = 1.0 / p.k_bolt / (p.temprescale * p.temperature);
   c.beta_m = 1.0 / p.k_bolt / (p.temprescale * p.temperature + 1.0e-15);
   auto bar_all = []() { __syncthreads(); };
   auto bar_draw = []() { asm volatile("bar.sync 2, 256;" ::: "memory"); };
   auto batch_range = [&](int gs, int ge, int& pb, int& pe) {
      pb = c.grec[(size_t)gs * c.gw];
      pe = (int)c.grec[(size_t)(ge - 1) * c.gw] + (int)c.grec[(size_t)(ge - 1) * c.gw + 1];
   };
   if (nph > 0) {
      int pb, pe;
      batch_range(gt[MCR_PH0] & 0x7fff, gt[MCR_PH0 + 2], pb, pe);
      mc_draw_batch<HB, NT>(t, p, c, tid, pb, pe, rec, rslot, blist, bcnt, B, bar_all);
   }
   mc_wait_tiles(tk, tile, c.k, tid);
   __syncthreads();
   mc_stage_list<NT, 4>(c, tid, cnt, ul);
   __syncthreads();
   // ---- batches: sweep warps on buffer (b & 1), draw warps fill buffer ((b + 1) & 1) ----
   int ph = 0, b = 0;
   while (ph < nph) {
      // phases [ph, pn) = batch b; the next batch starts at phase pn
      int pn = ph + 1;
      while (pn < nph && !(gt[MCR_PH0 + 4 * pn] & 0x8000u)) pn++;
      int pb, pe;
      batch_range(gt[MCR_PH0 + 4 * ph] & 0x7fff, gt[MCR_PH0 + 4 * ph + 2], pb, pe);
      if (drawer) {
         if (pn < nph) {
            int qb, qe;
            batch_range(gt[MCR_PH0 + 4 * pn] & 0x7fff, gt[MCR_PH0 + 4 * pn + 2], qb, qe);
            const int nb = (b + 1) & 1;
            mc_draw_batch<HB, 256>(t, p, c, tid - 256, qb, qe, rec + nb * RW * B, rslot + nb * B, blist, bcnt, B, bar_draw);
         }
      } else {
         const double* __restrict__ rb = rec + (b & 1) * RW * B;
         const unsigned short* __restrict__ sb = rslot + (b & 1) * B;
         for (int q = ph; q < pn; q++) {
            const uint2 pw = *reinterpret_cast<const uint2*>(gt + MCR_PH0 + 4 * q);
            const int ga = pw.x & 0x7fff, gb = pw.x >> 16;
            for (int gq = ga + warp; gq < gb; gq += 8) mc_sweep_group<HB>(t, p, mr, c, gq, lane, pb, rb, sb);
            asm volatile("bar.sync 1, 256;" ::: "memory");
         }
      }
      __syncthreads();      // batch switch: buffer (b + 1) & 1 complete, buffer b & 1 free
      ph = pn; b++;
   }
   __threadfence();
   __syncthreads();
   if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(tk.done + (size_t)c.k * tk.ntile + tile), "r"(tk.epoch) : "memory");
}

}  // namespace asd
