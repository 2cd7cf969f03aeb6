// Device code of the B200-native spin-dynamics hot path (sm_100a).
//
// Data layout in HBM:
//   * spins are packed as 32-byte vectors {ex, ey, ez, m} (unit direction + magnitude), one 256-bit
//     load (LDG.E.256) per gathered neighbour, two buffers per ensemble: `cur` (state at t) and `pred`
//     (midpoint / predictor state);
//   * atoms are stored in DEVICE ORDER: grouped by Hamiltonian index (aHam) and padded per group to a
//     multiple of 32, so that a warp always works on one sublattice: its coupling reads are broadcasts
//     and its neighbour gathers are unit-stride;
//   * neighbour tables are slot-major  nl[j][Npad]  (device indices), so a warp reads 128 contiguous bytes
//     per neighbour slot; reduced-Hamiltonian couplings are staged in shared memory once per CTA.
//
// What each routine restates (reference paths relative to the reference root):
//   site_field        source/Hamiltonian/hamiltonianactions.f90:185-243 (term order kept: Heisenberg, DM,
//                     BQ, anisotropy, external field; neighbour order j = 1..nlistsize kept)
//   midpoint stages   source/Evolution/midpoint.f90:123-178, :274-319
//   Depondt stages    source/Evolution/depondt.f90:136-190, :284-331
//   moment update     source/Evolution/updatemoments.f90:48-145
//   noise amplitude   source/RNG/randomnumbers.f90:667-670,735-746 (midpoint), depondt.f90:143-145
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace asd {

struct __align__(32) SpinVec {
   double x, y, z, m;
};

// Kernel parameter block (passed by value; < 4 KB).
struct Tables {
   int N;      // real atoms
   int Npad;   // device slots (groups padded to 32; in slab mode the neighbours' halo slots come last)
   int Nown;   // slots that are computed ( = Npad unless this engine holds a slab with halos)
   unsigned int atom_offset;  // global index of local atom 0 (slab decomposition): keys the noise, never the data
   unsigned int ens_offset;   // global index of local ensemble 0 (ensemble sharding): keys the noise
   int M;      // ensembles
   int NH;     // Hamiltonian rows
   int reduced;  // 1: couplings indexed by ham row (shared by a whole sublattice), 0: per atom
   const int* __restrict__ ham;   // [Npad] 0-based ham row of the slot, -1 for padding slots
   const int* __restrict__ orig;  // [Npad] 0-based original atom index, -1 for padding slots
   const int2* __restrict__ meta; // [Npad] {ham, orig} zipped: one 8-byte load in the stage kernels
   // Heisenberg (jtens = 1: tensorial exchange, nine couplings per pair, J(a,b) at component a + 3 b)
   int jtens;
   int z;
   const int* __restrict__ nl;      // [z][Npad] device index of neighbour
   const double* __restrict__ cp;   // reduced: [NH][z]; else [z][Npad]
   const int* __restrict__ lsize;   // reduced: [NH]; else unused (zero-padded couplings)
   // DM
   int zdm;
   const int* __restrict__ dml;     // [zdm][Npad]
   const double* __restrict__ dmv;  // reduced: [NH][zdm][3]; else [3][zdm][Npad]
   const int* __restrict__ dmsize;
   // BQ
   int zbq;
   const int* __restrict__ bql;
   const double* __restrict__ jbq;  // reduced: [NH][zbq]; else [zbq][Npad]
   const int* __restrict__ bqsize;
   // anisotropy (device order)
   int do_aniso;
   const int* __restrict__ taniso;     // [Npad]
   const double* __restrict__ eaniso;  // [3][Npad]
   const double* __restrict__ kaniso;  // [2][Npad]
   const double* __restrict__ sb;      // [Npad]
   int aniso_rows;                     // 1: anisotropy is the same for every atom of a Hamiltonian row -> aniso_small
   double aniso_small[8][8];           // [row]{taniso, k1, k2, ex, ey, ez, sb, -} in the kernel-parameter constant bank
   // external field: uniform vector or per-slot array [M][3][Npad]
   int ext_uniform;
   double hext[3];
   const double* __restrict__ ext;
   // spin-transfer torque field btorque [M][3][Npad] or null
   const double* __restrict__ btorque;
   // shared-memory staging of reduced couplings (doubles): cp | dmv | jbq
   int sm_cp, sm_dm, sm_bq;  // element counts (0 => read from global)
   // vectorised exchange table: nl4[zq][Npad] holds slots 4q..4q+3 of every atom (one 16-byte load per lane
   // per four neighbours); cp4[zq][Npad] the matching per-atom couplings (non-reduced only)
   int zq;
   const int4* __restrict__ nl4;
   const int* __restrict__ nlrow;   // [Npad][z] atom-major copy of nl (Monte Carlo layout): a sub-warp reads one atom's list coalesced
   const double* __restrict__ cprow; // [Npad][z] atom-major copy of the per-atom couplings (non-reduced Monte Carlo layout), zero-padded
   const double4* __restrict__ cp4;
   // staged tile gather: per 256-slot tile a sorted list (ulist) of the unique slots its atoms gather from; the
   // tile's CTA stages emomM of those slots in shared memory once, and nl16 holds the exchange neighbours as
   // 16-bit positions in that list (8 per 16-byte word) -- half the index bytes of nl, no global gathers
   int staged;            // 1: the LLG stage kernels use the tile path
   int ucap;              // row stride of ulist (largest unique count over the tiles, multiple of 32)
   const int* __restrict__ ulist;     // [ntile][ucap]
   const int* __restrict__ ucount;    // [ntile]
   const uint4* __restrict__ nl16;    // [zq8][Npad]
   int zq8;               // ceil(z / 8)
   const uint4* __restrict__ dm16;    // [ceil(zdm/8)][Npad] DM neighbours as positions in the tile's list (null: gather from global)
   const uint4* __restrict__ bq16;    // [ceil(zbq/8)][Npad] same for the biquadratic table
   // run-compressed table of lattice layouts (asd_runs.cuh): groups of `runs` x-runs share one union row
   int tile_slots;        // slots per tile of ulist / ucount (256; 512 or 1024 with the run kernel on super-bricks)
   int runs;              // 0: off, else 4: the LLG stage kernels use llg_runs_kernel (4 x-runs per warp)
   int urow;              // row stride of utab in 16-byte words
   const uint4* __restrict__ utab;    // [groups][urow]
   int union_max;   // largest number of distinct neighbour runs of a group (asd_layout_info)
   int mm;          // 1: the union rows carry 8 * base and the run kernels stage from the moment planes (MM instantiations)
   int pf_tiles;    // L2 bulk-prefetch distance in 256-atom tiles (0 = off)
   int cpl_param;   // 1: reduced exchange couplings live in cpl_small (kernel parameter = constant bank)
   double cpl_small[256];
};

#ifndef ASD_MINB
#define ASD_MINB 2
#endif
#ifndef ASD_CHUNK
#define ASD_CHUNK 8
#endif
#ifndef ASD_MINB_STAGED
#define ASD_MINB_STAGED 4   // CTAs per SM the staged stage kernels are compiled for (register budget 65536/(256*n))
#endif

struct LlgParams {
   int per_site;        // 1: read Landeg/lambda/Temp arrays (device order), 0: uniform scalars
   double landeg, lambda, temp;
   const double* __restrict__ landeg_a;
   const double* __restrict__ lambda_a;
   const double* __restrict__ temp_a;
   double delta_t, gamma, k_bolt, mub, temprescale;
   int mompar;
   const double* __restrict__ mmom0;  // [M][Npad], only read when mompar != 0
   unsigned long long seed;
   unsigned long long step;  // value of mstep for this step (keys the noise)
   int thermal;              // 0: skip noise entirely
   // uniform case (per_site == 0): site-independent factors evaluated once on the host with the same IEEE
   // operations the kernel would use (division and square root are correctly rounded on both sides)
   double u_lldamp, u_dt, u_sqrtdt, u_Dk, u_Dp;
   // thermal amplitude without its 1/sqrt(m) factor: sigma = u_sig * m^-1/2 (u_sig1: midpoint, sqrt(2 u_Dk T temprescale);
   // u_sig5: Depondt, sqrt(u_Dp T temprescale))
   double u_sig1, u_sig5;
   // fused observable: when non-null, the corrector launch also leaves sum_i emomM of every tile in
   // msum_part[k][ntile][4] (asd_measure then only adds the per-tile partials: no second pass over the spins)
   double* msum_part;
   int msum_ntile;
   // fixed-moment run (Nred < Natom, red_atom_list of evolve_first, evolution.f90:38-44): frozen[slot] != 0 marks an atom
   // that is NOT in the list of evolving atoms -- the integrators skip it, every neighbour still sees its moment
   const unsigned char* __restrict__ frozen;
   // time-dependent uniform field (asd_set_time_field; time_external_field of hamiltonianactions.f90:241, global part of
   // calculatefields.f90:92-185): tf = the vector of THIS step (stage launches: set by the host per step, zero without a schedule,
   // added unconditionally -- a branch here cost the register-blocked run kernel 18 %, measured); tfield[step - tf_first][3] = the
   // device copy of the schedule for the resident kernel, which advances the step itself
   double tf[3];
   // moment planes of the MM run kernels (asd_runs.cuh): emomM of `cur` / `pred` as [M][3][Npad], null on every other path
   double* mm_cur;
   double* mm_pred;
   const double* __restrict__ tfield;
   long long tf_first;
   int tf_n;
};

// resident kernel: the schedule entry of `step`
__device__ __forceinline__ void add_time_field(const LlgParams& p, unsigned long long step, double h[3]) {
   if (p.tfield) {
      const long long q = (long long)step - p.tf_first;
      if (q >= 0 && q < p.tf_n) {
         const double* __restrict__ f = p.tfield + (size_t)q * 3;
         h[0] += __ldg(f); h[1] += __ldg(f + 1); h[2] += __ldg(f + 2);
      }
   }
}

// ------------------------------------------------------------------------------------------------
// Counter-based Gaussian noise: Philox4x32-10 (Salmon et al. 2011) keyed by the run seed; the counter is
// (original atom index, ensemble, step, draw) so the stream is independent of device order, of the slab /
// ensemble decomposition and of the GPU count.  Two calls give four 64-bit words -> two Box-Muller pairs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
   const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
   for (int r = 0; r < 10; r++) {
      uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
      uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
   }
   out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, uint32_t c, uint32_t d, double& g0, double& g1) {
   const uint64_t w0 = ((uint64_t)a << 32) | b, w1 = ((uint64_t)c << 32) | d;
   const double u1 = (double)((w0 >> 11) + 1ull) * (1.0 / 9007199254740992.0);  // (0,1]
   const double u2 = (double)(w1 >> 11) * (1.0 / 9007199254740992.0);           // [0,1)
   const double rad = sqrt(-2.0 * log(u1));
   double s, c2;
   sincospi(2.0 * u2, &s, &c2);
   g0 = rad * c2;
   g1 = rad * s;
}

// three N(0,1) numbers for (atom, ensemble, step); `stream` separates LLG noise from MC draws.
__device__ __forceinline__ void gauss3(unsigned long long seed, uint32_t atom, uint32_t ens, unsigned long long step,
                                       uint32_t stream, double& g0, double& g1, double& g2) {
   uint32_t r[4];
   const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
   const uint32_t s_lo = (uint32_t)step, s_hi = (uint32_t)(step >> 32) ^ (stream << 24);
   philox4x32_10(atom, ens, s_lo, s_hi, k0, k1, r);
   box_muller(r[0], r[1], r[2], r[3], g0, g1);
   double dummy;
   philox4x32_10(atom, ens | 0x80000000u, s_lo, s_hi, k0, k1, r);
   box_muller(r[0], r[1], r[2], r[3], g2, dummy);
}

// Langevin noise of the LLG stages: three N(0,1) numbers from ONE Philox call, Box-Muller evaluated in single
// precision (accurate logf, MUFU sin/cos on [-pi,pi)) and widened to FP64.  The reference's own Gaussian source
// is the single-precision-resolution Ziggurat r4_nor (source/RNG/randomnumbers.f90:330-438: a 32-bit integer hz
// times a table entry), so 32-bit-resolution variates are what the Fortran path feeds into the same formulas.
// Uniforms use all 32 bits (u = r 2^-32 + 2^-33, tail to 6.7 sigma).  ~5x fewer instructions than the FP64 form.
__device__ __forceinline__ void gauss3f_raw(unsigned long long seed, uint32_t atom, uint32_t ens, unsigned long long step,
                                            uint32_t stream, float& g0, float& g1, float& g2) {
   uint32_t r[4];
   const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
   const uint32_t s_lo = (uint32_t)step, s_hi = (uint32_t)(step >> 32) ^ (stream << 24);
   philox4x32_10(atom, ens, s_lo, s_hi, k0, k1, r);
   const float two_m32 = 2.3283064365386963e-10f, two_m33 = 1.1641532182693481e-10f;
   const float u1 = fmaf((float)r[0], two_m32, two_m33), u3 = fmaf((float)r[2], two_m32, two_m33);
   const float a1 = ((float)r[1] * two_m32 - 0.5f) * 6.283185307179586f;
   const float a2 = ((float)r[3] * two_m32 - 0.5f) * 6.283185307179586f;
   const float rad1 = sqrtf(fmaxf(-2.0f * logf(u1), 0.0f)), rad2 = sqrtf(fmaxf(-2.0f * logf(u3), 0.0f));
   float s1, c1, s2, c2;
   __sincosf(a1, &s1, &c1);
   __sincosf(a2, &s2, &c2);
   (void)s2;
   g0 = rad1 * c1;
   g1 = rad1 * s1;
   g2 = rad2 * c2;
}
__device__ __forceinline__ void gauss3f(unsigned long long seed, uint32_t atom, uint32_t ens, unsigned long long step,
                                        uint32_t stream, double& g0, double& g1, double& g2) {
   float a, b, c;
   gauss3f_raw(seed, atom, ens, step, stream, a, b, c);
   g0 = (double)a; g1 = (double)b; g2 = (double)c;
}

// four U[0,1) numbers (53-bit) for Monte Carlo draws
__device__ __forceinline__ void uniform4(unsigned long long seed, uint32_t atom, uint32_t ens, unsigned long long step,
                                         uint32_t stream, double* u) {
   uint32_t r[4];
   const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
   const uint32_t s_lo = (uint32_t)step, s_hi = (uint32_t)(step >> 32) ^ (stream << 24);
   philox4x32_10(atom, ens, s_lo, s_hi, k0, k1, r);
   u[0] = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
   u[1] = (double)((((uint64_t)r[2] << 32) | r[3]) >> 11) * (1.0 / 9007199254740992.0);
   philox4x32_10(atom, ens | 0x80000000u, s_lo, s_hi, k0, k1, r);
   u[2] = (double)((((uint64_t)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
   u[3] = (double)((((uint64_t)r[2] << 32) | r[3]) >> 11) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// Effective field of one site.  S = spin buffer of this ensemble (device order).  Returns the bilinear
// part (bs: exchange, DM, uniaxial) and the "q" part (bq: biquadratic, cubic part of taniso 7), like
// beff_s / beff_q in hamiltonianactions.f90:185-238.  own = this site's packed spin in S.
// smc/smd/smb: shared-memory copies of the reduced couplings (or null -> global).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 ld_nc_d4(const double4* p) {
   double4 v;
   asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
   return v;
}

// Heisenberg sum with explicit memory-level parallelism (hamiltonianactions.f90:461-464, same j order):
// neighbour indices arrive as 16-byte vectors, one chunk (ASD_CHUNK slots) ahead of the gathers that use
// them; the ASD_CHUNK 256-bit gathers of a chunk are all issued before the first FMA consumes one.
template <bool REDUCED, int CH = ASD_CHUNK>
__device__ __forceinline__ void exchange_chunked(const Tables& t, const SpinVec* __restrict__ S, int i, int ih,
                                                 const double* smc, double& fx, double& fy, double& fz) {
   constexpr int CQ = CH / 4;
   const size_t Npad = t.Npad;
   const int n = REDUCED ? __ldg(t.lsize + ih) : t.z;
   const int nq = (n + 3) >> 2;
   const int4* __restrict__ p = t.nl4 + i;
   const double4* __restrict__ pc = REDUCED ? nullptr : t.cp4 + i;
   const double* __restrict__ crow = REDUCED ? (smc ? smc + (size_t)ih * t.z : t.cp + (size_t)ih * t.z) : nullptr;
   const int cbase = ih * t.z;
   int4 cur[CQ], nxt[CQ];
   double4 ccur[REDUCED ? 1 : CQ], cnxt[REDUCED ? 1 : CQ];
#pragma unroll
   for (int q = 0; q < CQ; q++) {
      if (q < nq) { cur[q] = __ldg(p + q * Npad); if (!REDUCED) ccur[q] = ld_nc_d4(pc + q * Npad); }
   }
   for (int j0 = 0; j0 < n; j0 += CH) {
      const int q1 = (j0 + CH) >> 2;
#pragma unroll
      for (int q = 0; q < CQ; q++)
         if (q1 + q < nq) { nxt[q] = __ldg(p + (size_t)(q1 + q) * Npad); if (!REDUCED) cnxt[q] = ld_nc_d4(pc + (size_t)(q1 + q) * Npad); }
      SpinVec v[CH];
#pragma unroll
      for (int u = 0; u < CH; u++) {
         const int4 w = cur[u >> 2];
         const int nb = (u & 3) == 0 ? w.x : (u & 3) == 1 ? w.y : (u & 3) == 2 ? w.z : w.w;
         if (j0 + u < n) v[u] = S[nb];
      }
#pragma unroll
      for (int u = 0; u < CH; u++) {
         if (j0 + u < n) {
            double cj;
            if (REDUCED) cj = t.cpl_param ? t.cpl_small[cbase + j0 + u] : crow[j0 + u];
            else { const double4 w = ccur[u >> 2]; cj = (u & 3) == 0 ? w.x : (u & 3) == 1 ? w.y : (u & 3) == 2 ? w.z : w.w; }
            fx = fma(cj, v[u].x * v[u].m, fx);
            fy = fma(cj, v[u].y * v[u].m, fy);
            fz = fma(cj, v[u].z * v[u].m, fz);
         }
      }
#pragma unroll
      for (int q = 0; q < CQ; q++) { cur[q] = nxt[q]; if (!REDUCED) ccur[q] = cnxt[q]; }
   }
}

// streaming 16-byte load of the index table: read-only path, do not allocate in L1 (the words are used once)
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
   uint4 v;
   asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
   return v;
}

#ifndef ASD_NPF
#define ASD_NPF 2   // index words (8 neighbours each) kept in flight per thread
#endif

// first ASD_NPF index words of atom i (issued before the tile is staged so that they overlap the staging)
__device__ __forceinline__ void idx_prologue(const Tables& t, int i, int nq, uint4 w[ASD_NPF]) {
   const uint4* __restrict__ p = t.nl16 + i;
#pragma unroll
   for (int q = 0; q < ASD_NPF; q++)
      if (q < nq) w[q] = ld_stream_u4(p + (size_t)q * t.Npad);
}

// Heisenberg sum from the staged tile (hamiltonianactions.f90:461-464, same j order): neighbours arrive as 16-bit
// positions into the CTA's shared-memory copy of emomM (s3: 24-byte records {mx,my,mz}; the 24-byte stride is
// conflict-free for 8-byte accesses), three LDS.64 per neighbour off one address.
// CSRC: 0 = reduced couplings in the kernel-parameter constant bank, 1 = reduced couplings through a pointer
// (shared or global), 2 = per-atom couplings streamed from cp4.
template <int CSRC>
__device__ __forceinline__ void exchange_staged_impl(const Tables& t, const double* __restrict__ s3, int i, int ih, int n,
                                                     int nq, uint4 w[ASD_NPF], const double* __restrict__ crow,
                                                     double& fx, double& fy, double& fz) {
   const size_t Npad = t.Npad;
   const uint4* __restrict__ p = t.nl16 + i;
   const double4* __restrict__ pc = (CSRC == 2) ? t.cp4 + i : nullptr;
   const int cbase = ih * t.z;
   const int nfull = n >> 3;
   for (int q0 = 0; q0 < nq; q0 += ASD_NPF) {
#pragma unroll
      for (int s = 0; s < ASD_NPF; s++) {
         const int q = q0 + s;
         if (q < nq) {
            const uint4 c = w[s];
            if (q + ASD_NPF < nq) w[s] = ld_stream_u4(p + (size_t)(q + ASD_NPF) * Npad);
            double cj[8];
            if (CSRC == 2) {
               const double4 ca = ld_nc_d4(pc + (size_t)(2 * q) * Npad);
               double4 cb = make_double4(0.0, 0.0, 0.0, 0.0);
               if (8 * q + 4 < n) cb = ld_nc_d4(pc + (size_t)(2 * q + 1) * Npad);
               cj[0] = ca.x; cj[1] = ca.y; cj[2] = ca.z; cj[3] = ca.w; cj[4] = cb.x; cj[5] = cb.y; cj[6] = cb.z; cj[7] = cb.w;
            }
            const unsigned li[8] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu, c.y >> 16,
                                    c.z & 0xffffu, c.z >> 16, c.w & 0xffffu, c.w >> 16};
            if (q < nfull) {
               // full word: eight neighbours, no predicates
#pragma unroll
               for (int u = 0; u < 8; u++) {
                  const double* __restrict__ m = s3 + li[u] * 3u;
                  const double cc = (CSRC == 0) ? t.cpl_small[cbase + 8 * q + u] : (CSRC == 1) ? crow[8 * q + u] : cj[u];
                  fx = fma(cc, m[0], fx);
                  fy = fma(cc, m[1], fy);
                  fz = fma(cc, m[2], fz);
               }
            } else {
#pragma unroll
               for (int u = 0; u < 8; u++) {
                  if (8 * q + u < n) {
                     const double* __restrict__ m = s3 + li[u] * 3u;
                     const double cc = (CSRC == 0) ? t.cpl_small[cbase + 8 * q + u] : (CSRC == 1) ? crow[8 * q + u] : cj[u];
                     fx = fma(cc, m[0], fx);
                     fy = fma(cc, m[1], fy);
                     fz = fma(cc, m[2], fz);
                  }
               }
            }
         }
      }
   }
}

template <bool REDUCED>
__device__ __forceinline__ void exchange_staged(const Tables& t, const double* __restrict__ s3, int i, int ih, int n, int nq,
                                                uint4 w[ASD_NPF], const double* smc, double& fx, double& fy, double& fz) {
   if (!REDUCED) exchange_staged_impl<2>(t, s3, i, ih, n, nq, w, nullptr, fx, fy, fz);
   else if (t.cpl_param) exchange_staged_impl<0>(t, s3, i, ih, n, nq, w, nullptr, fx, fy, fz);
   else exchange_staged_impl<1>(t, s3, i, ih, n, nq, w, smc ? smc + (size_t)ih * t.z : t.cp + (size_t)ih * t.z, fx, fy, fz);
}

// L2 bulk prefetch of the index words and gather list of the tile `pf_tiles` ahead (staged path)
__device__ __forceinline__ void prefetch_tile_staged(const Tables& t, int this_tile) {
   if (t.pf_tiles == 0) return;
   const size_t tile = (size_t)this_tile + t.pf_tiles;
   const size_t first = tile * blockDim.x;
   if (first >= (size_t)t.Nown) return;
   const unsigned cnt = (unsigned)min((size_t)blockDim.x, (size_t)t.Nown - first);
   if ((int)threadIdx.x < t.zq8) {
      const uint4* a = t.nl16 + (size_t)threadIdx.x * t.Npad + first;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(cnt * 16u) : "memory");
   } else if ((int)threadIdx.x == t.zq8) {
      const int* a = t.ulist + tile * t.ucap;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)t.ucap * 4u) : "memory");
   }
}

// L2 bulk prefetch (cp.async.bulk.prefetch.L2) of the index rows and spins of the tile that a CTA scheduled
// ~one wave later will work on: turns the DRAM latency of the index stream into an L2 hit.
__device__ __forceinline__ void prefetch_tile(const Tables& t, const SpinVec* S, int this_tile) {
   if (t.pf_tiles == 0 || t.nl4 == nullptr) return;
   const size_t tile = (size_t)this_tile + t.pf_tiles;
   const size_t first = tile * blockDim.x;
   if (first >= (size_t)t.Nown) return;
   const unsigned cnt = (unsigned)min((size_t)blockDim.x, (size_t)t.Nown - first);
   if ((int)threadIdx.x < t.zq) {
      const int4* a = t.nl4 + (size_t)threadIdx.x * t.Npad + first;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(cnt * 16u) : "memory");
   } else if ((int)threadIdx.x == t.zq) {
      const SpinVec* a = S + first;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(cnt * 32u) : "memory");
   }
}

// EXCH = false: the Heisenberg sum was already accumulated into bs[] by the caller (staged tile path).
// 16-bit position j of a neighbour word row (8 positions per 16-byte word)
__device__ __forceinline__ unsigned pos16(const uint4* __restrict__ tab, size_t Npad, int i, int j) {
   const uint4 w = __ldg(tab + (size_t)(j >> 3) * Npad + i);
   const unsigned c = ((j >> 1) & 3) == 0 ? w.x : ((j >> 1) & 3) == 1 ? w.y : ((j >> 1) & 3) == 2 ? w.z : w.w;
   return (j & 1) ? (c >> 16) : (c & 0xffffu);
}

// s3: the CTA's shared-memory copy of emomM of the tile's gather list (staged kernels), or null.  With s3 and the
// 16-bit position tables dm16 / bq16 the DM and BQ neighbours are read from shared memory too.
// XS (compile time: a kernel without DM / BQ work must not carry this code, it costs the plain Heisenberg kernel 15 %).
// one neighbour of the DM sum (hamiltonianactions.f90:566-571) and of the biquadratic sum (:791-794); shared by every
// kernel that walks these lists so that the arithmetic is the same instruction sequence everywhere
__device__ __forceinline__ void dm_term(double Dx, double Dy, double Dz, double mx, double my, double mz, double& fx, double& fy,
                                        double& fz) {
   fx = fx + Dz * my - Dy * mz;
   fy = fy + Dx * mz - Dz * mx;
   fz = fz + Dy * mx - Dx * my;
}
__device__ __forceinline__ void bq_term(double jb, double mx, double my, double mz, double ox, double oy, double oz, double& qx,
                                        double& qy, double& qz) {
   const double dot = mx * ox + my * oy + mz * oz;
   const double c = 2.0 * jb * dot;
   qx = fma(c, mx, qx);
   qy = fma(c, my, qy);
   qz = fma(c, mz, qz);
}

// single-ion anisotropy field of slot i at the moment (ox, oy, oz) (hamiltonianactions.f90:225-239, 842-919): the uniaxial
// part and the cubic part of taniso 2 join the bilinear field f, the cubic part of taniso 7 (times sb) the "q" field
template <bool REDUCED>
__device__ __forceinline__ void aniso_field(const Tables& t, int i, int ih, double ox, double oy, double oz, double& fx, double& fy,
                                            double& fz, double& qx, double& qy, double& qz) {
   const int Npad = t.Npad;
   if (t.do_aniso) {
      const bool rows = REDUCED && t.aniso_rows;
      const int ta = rows ? (int)t.aniso_small[ih][0] : __ldg(t.taniso + i);
      if (ta == 1 || ta == 2 || ta == 7) {
         const double k1 = rows ? t.aniso_small[ih][1] : __ldg(t.kaniso + i), k2 = rows ? t.aniso_small[ih][2] : __ldg(t.kaniso + Npad + i);
         if (ta == 1 || ta == 7) {
            const double ex = rows ? t.aniso_small[ih][3] : __ldg(t.eaniso + i), ey = rows ? t.aniso_small[ih][4] : __ldg(t.eaniso + Npad + i),
                         ez = rows ? t.aniso_small[ih][5] : __ldg(t.eaniso + 2 * (size_t)Npad + i);
            const double tt1 = ox * ex + oy * ey + oz * ez;
            const double tt2 = k1 + 2.0 * k2 * (1.0 - tt1 * tt1);
            const double tt3 = 2.0 * tt1 * tt2;
            fx -= tt3 * ex; fy -= tt3 * ey; fz -= tt3 * ez;
         }
         if (ta == 2 || ta == 7) {
            const double x2 = ox * ox, y2 = oy * oy, z2 = oz * oz;
            const double cx = 2.0 * k1 * ox * (y2 + z2) + 2.0 * k2 * ox * (y2 * z2);
            const double cy = 2.0 * k1 * oy * (z2 + x2) + 2.0 * k2 * oy * (z2 * x2);
            const double cz = 2.0 * k1 * oz * (x2 + y2) + 2.0 * k2 * oz * (x2 * y2);
            if (ta == 2) { fx += cx; fy += cy; fz += cz; }
            else { const double s = rows ? t.aniso_small[ih][6] : __ldg(t.sb + i); qx += cx * s; qy += cy * s; qz += cz * s; }
         }
      }
   }
}

// PAIRS = false: every pair sum (Heisenberg, DM, BQ) was accumulated by the caller into bs[] / bq[] (resident kernel).
template <bool REDUCED, bool EXCH = true, int CH = ASD_CHUNK, bool XS = false, bool PAIRS = true>
__device__ __forceinline__ void site_field(const Tables& t, const SpinVec* __restrict__ S, int i, int ih,
                                           const SpinVec& own, const double* smc, const double* smd,
                                           const double* smb, double bs[3], double bq[3], const double* __restrict__ s3 = nullptr,
                                           const uint4* dmw0 = nullptr) {
   // dmw0: the first DM position word of this atom when the caller fetched it ahead of time (asd_runs.cuh)
   double fx = EXCH ? 0.0 : bs[0], fy = EXCH ? 0.0 : bs[1], fz = EXCH ? 0.0 : bs[2];
   const int Npad = t.Npad;
   // ---- Heisenberg (hamiltonianactions.f90:461-464) ----
   if (!EXCH) {}
   else if (t.nl4) exchange_chunked<REDUCED, CH>(t, S, i, ih, smc, fx, fy, fz);
   else if (t.jtens) {
      // tensor_field (hamiltonianactions.f90:499-542): f += J(:,1) m_x + J(:,2) m_y + J(:,3) m_z
      const int* __restrict__ nl = t.nl + i;
      const int n = REDUCED ? __ldg(t.lsize + ih) : t.z;
      for (int j = 0; j < n; j++) {
         const SpinVec v = S[__ldg(nl + (size_t)j * Npad)];
         const double mx = v.x * v.m, my = v.y * v.m, mz = v.z * v.m;
         double J[9];
         if (REDUCED) {
            const double* __restrict__ c = (smc ? smc : t.cp) + ((size_t)ih * t.z + j) * 9;
#pragma unroll
            for (int a = 0; a < 9; a++) J[a] = c[a];
         } else {
#pragma unroll
            for (int a = 0; a < 9; a++) J[a] = __ldg(t.cp + ((size_t)a * t.z + j) * Npad + i);
         }
         fx = fx + J[0] * mx + J[3] * my + J[6] * mz;
         fy = fy + J[1] * mx + J[4] * my + J[7] * mz;
         fz = fz + J[2] * mx + J[5] * my + J[8] * mz;
      }
   } else {
      const int* __restrict__ nl = t.nl + i;
      if (REDUCED) {
         const int n = __ldg(t.lsize + ih);
         const double* __restrict__ c = smc ? smc + (size_t)ih * t.z : t.cp + (size_t)ih * t.z;
#pragma unroll 10
         for (int j = 0; j < n; j++) {
            const int nb = __ldg(nl + (size_t)j * Npad);
            const SpinVec v = S[nb];
            const double cj = c[j];
            fx = fma(cj, v.x * v.m, fx);
            fy = fma(cj, v.y * v.m, fy);
            fz = fma(cj, v.z * v.m, fz);
         }
      } else {
         const double* __restrict__ c = t.cp + i;
#pragma unroll 10
         for (int j = 0; j < t.z; j++) {
            const int nb = __ldg(nl + (size_t)j * Npad);
            const double cj = __ldg(c + (size_t)j * Npad);
            const SpinVec v = S[nb];
            fx = fma(cj, v.x * v.m, fx);
            fy = fma(cj, v.y * v.m, fy);
            fz = fma(cj, v.z * v.m, fz);
         }
      }
   }
   // ---- Dzyaloshinskii-Moriya (hamiltonianactions.f90:565-571) ----
   if (PAIRS && t.zdm > 0) {
      const int* __restrict__ nl = t.dml + i;
      const int n = REDUCED ? __ldg(t.dmsize + ih) : t.zdm;
      const bool smem_dm = XS && t.dm16 != nullptr;
      uint4 wd = make_uint4(0u, 0u, 0u, 0u);
      int jstart = 0;
      if (XS && REDUCED && smem_dm && smd != nullptr && dmw0 != nullptr) {
         // register-blocked kernel: the first eight DM neighbours from the prefetched position word, D vectors and moments from
         // shared memory, fully unrolled (the generic loop below costs 42 instructions per neighbour: variable shifts to extract the
         // position, generic loads of D, a dependent global load of the word; ncu r2q: a quarter of the instructions of config 4)
         __builtin_assume(__isShared(smd));
         const uint4 w = *dmw0;
         const unsigned li[8] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16, w.z & 0xffffu, w.z >> 16, w.w & 0xffffu, w.w >> 16};
         const double* __restrict__ D0 = smd + (size_t)ih * t.zdm * 3;
#pragma unroll
         for (int j = 0; j < 8; j++)
            if (j < n) {
               const double* __restrict__ m = s3 + li[j] * 3u;
               dm_term(D0[3 * j], D0[3 * j + 1], D0[3 * j + 2], m[0], m[1], m[2], fx, fy, fz);
            }
         jstart = 8;
      }
      for (int j = jstart; j < n; j++) {
         double Dx, Dy, Dz;
         if (REDUCED) {
            const double* __restrict__ d = (smd ? smd : t.dmv) + ((size_t)ih * t.zdm + j) * 3;
            Dx = d[0]; Dy = d[1]; Dz = d[2];
         } else {
            const size_t o = (size_t)j * Npad + i, s = (size_t)t.zdm * Npad;
            Dx = __ldg(t.dmv + o); Dy = __ldg(t.dmv + s + o); Dz = __ldg(t.dmv + 2 * s + o);
         }
         double mx, my, mz;
         if (smem_dm) {
            if ((j & 7) == 0) wd = (j == 0 && dmw0) ? *dmw0 : __ldg(t.dm16 + (size_t)(j >> 3) * Npad + i);
            const unsigned c = ((j >> 1) & 3) == 0 ? wd.x : ((j >> 1) & 3) == 1 ? wd.y : ((j >> 1) & 3) == 2 ? wd.z : wd.w;
            const double* __restrict__ m = s3 + ((j & 1) ? (c >> 16) : (c & 0xffffu)) * 3u;
            mx = m[0]; my = m[1]; mz = m[2];
         } else {
            const SpinVec v = S[__ldg(nl + (size_t)j * Npad)];
            mx = v.x * v.m; my = v.y * v.m; mz = v.z * v.m;
         }
         dm_term(Dx, Dy, Dz, mx, my, mz, fx, fy, fz);
      }
   }
   double qx = PAIRS ? 0.0 : bq[0], qy = PAIRS ? 0.0 : bq[1], qz = PAIRS ? 0.0 : bq[2];
   const double ox = own.x * own.m, oy = own.y * own.m, oz = own.z * own.m;  // emomM of this site
   // ---- biquadratic (hamiltonianactions.f90:790-795) ----
   if (PAIRS && t.zbq > 0) {
      const int* __restrict__ nl = t.bql + i;
      const int n = REDUCED ? __ldg(t.bqsize + ih) : t.zbq;
      const bool smem_bq = XS && t.bq16 != nullptr;
      for (int j = 0; j < n; j++) {
         const double jb = REDUCED ? (smb ? smb : t.jbq)[(size_t)ih * t.zbq + j] : __ldg(t.jbq + (size_t)j * Npad + i);
         double mx, my, mz;
         if (smem_bq) {
            const double* __restrict__ m = s3 + pos16(t.bq16, Npad, i, j) * 3u;
            mx = m[0]; my = m[1]; mz = m[2];
         } else {
            const SpinVec v = S[__ldg(nl + (size_t)j * Npad)];
            mx = v.x * v.m; my = v.y * v.m; mz = v.z * v.m;
         }
         bq_term(jb, mx, my, mz, ox, oy, oz, qx, qy, qz);
      }
   }
   // ---- single-ion anisotropy (hamiltonianactions.f90:225-239, 842-919); uses the FULL moment ----
   aniso_field<REDUCED>(t, i, ih, ox, oy, oz, fx, fy, fz, qx, qy, qz);
   bs[0] = fx; bs[1] = fy; bs[2] = fz;
   bq[0] = qx; bq[1] = qy; bq[2] = qz;
}

// external field of slot i, ensemble k (external_field + time_external_field(=0), hamiltonianactions.f90:241)
__device__ __forceinline__ void ext_field(const Tables& t, int i, int k, double h[3]) {
   if (t.ext_uniform) { h[0] = t.hext[0]; h[1] = t.hext[1]; h[2] = t.hext[2]; }
   else {
      const double* __restrict__ p = t.ext + (size_t)k * 3 * t.Npad + i;
      h[0] = __ldg(p); h[1] = __ldg(p + t.Npad); h[2] = __ldg(p + 2 * (size_t)t.Npad);
   }
}

// stage the reduced couplings into shared memory (once per CTA); returns pointers (null => use global)
__device__ __forceinline__ void stage_couplings(const Tables& t, double* sm, const double*& smc, const double*& smd,
                                                const double*& smb) {
   smc = smd = smb = nullptr;
   const int n0 = t.cpl_param ? 0 : t.sm_cp, n1 = t.sm_dm, n2 = t.sm_bq;
   if (n0 + n1 + n2 == 0) return;
   for (int q = threadIdx.x; q < n0; q += blockDim.x) sm[q] = t.cp[q];
   for (int q = threadIdx.x; q < n1; q += blockDim.x) sm[n0 + q] = t.dmv[q];
   for (int q = threadIdx.x; q < n2; q += blockDim.x) sm[n0 + n1 + q] = t.jbq[q];
   __syncthreads();
   if (n0) smc = sm;
   if (n1) smd = sm + n0;
   if (n2) smb = sm + n0 + n1;
}

// m^-1/2 for the thermal amplitude (sigma ~ sqrt(T / m), randomnumbers.f90:667-670): single-precision seed + two Newton steps in
// double (relative error < 4e-16), no division, no square root, no slow-path branch -- the division 1/m and the square root of
// the literal formula cost the register-blocked kernel 45 instructions per atom-stage in its serial part.  m > 0.
__device__ __forceinline__ double rsqrt_nr(double m) {
   double y = (double)rsqrtf((float)m);
   const double hm = 0.5 * m;
   y = fma(y, fma(-hm * y, y, 0.5), y);
   y = fma(y, fma(-hm * y, y, 0.5), y);
   return y;
}

// Cayley transform of the semi-implicit midpoint scheme (midpoint.f90:153-164): (I+skew A)^-1 (I+skew A)^T e
__device__ __forceinline__ void cayley(const double e[3], const double A[3], double out[3]) {
   const double detAi = 1.0 / (1.0 + (A[0] * A[0] + A[1] * A[1] + A[2] * A[2]));
   const double a0 = e[0] + e[1] * A[2] - e[2] * A[1];
   const double a1 = e[1] + e[2] * A[0] - e[0] * A[2];
   const double a2 = e[2] + e[0] * A[1] - e[1] * A[0];
   const double t0 = a0 * (1 + A[0] * A[0]) + a1 * (A[0] * A[1] + A[2]) + a2 * (A[0] * A[2] - A[1]);
   const double t1 = a0 * (A[1] * A[0] - A[2]) + a1 * (1 + A[1] * A[1]) + a2 * (A[1] * A[2] + A[0]);
   const double t2 = a0 * (A[2] * A[0] + A[1]) + a1 * (A[2] * A[1] - A[0]) + a2 * (1 + A[2] * A[2]);
   out[0] = t0 * detAi; out[1] = t1 * detAi; out[2] = t2 * detAi;
}

// Rodrigues rotation of the Depondt scheme (depondt.f90:157-179)
__device__ __forceinline__ void rodrigues(const double bd[3], const double e[3], double dtg_lldamp, double out[3]) {
   double Bnorm = sqrt(bd[0] * bd[0] + bd[1] * bd[1] + bd[2] * bd[2]) + 1.0e-15;
   // one division, three products (depondt.f90:160-162 divides three times: the quotients differ in the last bit at most, inside the
   // 1e-12 parity bar; two FP64 divisions with their slow-path branches less per atom and stage)
   const double binv = 1.0 / Bnorm;
   const double hx = bd[0] * binv, hy = bd[1] * binv, hz = bd[2] * binv;
   const double v = Bnorm * dtg_lldamp;
   double sinv, cosv;
   sincos(v, &sinv, &cosv);
   const double u = 1.0 - cosv;
   out[0] = hx * hx * u * e[0] + cosv * e[0] + hx * hy * u * e[1] - hz * sinv * e[1] + hx * hz * u * e[2] + hy * sinv * e[2];
   out[1] = hy * hx * u * e[0] + hz * sinv * e[0] + hy * hy * u * e[1] + cosv * e[1] + hy * hz * u * e[2] - hx * sinv * e[2];
   out[2] = hx * hz * u * e[0] - hy * sinv * e[0] + hz * hy * u * e[1] + hx * sinv * e[1] + hz * hz * u * e[2] + cosv * e[2];
}

// moment magnitude update (updatemoments.f90:105-145)
__device__ __forceinline__ double calcm(int mompar, double m, double m0, double ez) {
   if (mompar == 1) return fmax(m0 * fabs(ez), 1e-4);
   if (mompar == 2) return fmax(m0 * (ez * ez), 1.0e-4);
   return m;
}

// ------------------------------------------------------------------------------------------------
// One stage of one LLG step, field evaluation fused with the integrator.
//   SOLVER 1 = semi-implicit midpoint, 5 = Depondt.   STAGE 1 = predictor, 2 = corrector (+ moment update).
//   STAGE 1: gathers from `cur`, writes `pred` (midpoint spin for SOLVER 1, rotated spin for SOLVER 5);
//   STAGE 2: gathers from `pred`, reads own `cur`, writes the new spin to `cur` (own slot only).
//   b2eff (Depondt only): [M][3][Npad] predictor field kept for the Heun average.
// ------------------------------------------------------------------------------------------------
// Halo push of the slab decomposition (one slab of the supercell per GPU): an EDGE launch covers the
// tiles that hold the H boundary planes of this slab; every atom of those planes also stores its new spin straight
// into the halo slots of the ring neighbour that gathers it (peer-mapped memory over NVLink), and the last CTA of
// the launch publishes the exchange epoch in the neighbours' flag words.  No separate pack / send / unpack passes.
struct EdgeParams {
   const int* __restrict__ hdst_lo;   // [Nown] halo slot in the LOWER neighbour that mirrors this atom, or -1
   const int* __restrict__ hdst_hi;   // [Nown] halo slot in the UPPER neighbour, or -1
   SpinVec* peer_lo;                  // the buffer this stage writes (cur or pred), as mapped from the neighbours
   SpinVec* peer_hi;
   double* peer_mlo;                  // moment planes of that buffer in the neighbours (MM run kernels), [M][3][Npad]
   double* peer_mhi;
   unsigned long long* flag_lo;       // flag word of the lower neighbour that THIS rank owns (its "upper" word)
   unsigned long long* flag_hi;
   unsigned long long epoch;          // value to publish
   unsigned int* ctr;                 // CTA completion counter (self-resetting)
};

// which tiles a launch covers: blockIdx.x < split -> first + blockIdx.x, else second + (blockIdx.x - split)
struct TileRange { int first, split, second; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

// waits until both neighbours have published `epoch` (their halo stores are then visible); gives up after
// `timeout` clock ticks and raises *err so that a lost peer cannot hang the GPU
__global__ void halo_wait_kernel(const unsigned long long* flags, int wait_lo, int wait_hi, unsigned long long epoch,
                                 long long timeout, int* err) {
   const long long t0 = clock64();
   for (int side = 0; side < 2; side++) {
      if (!(side == 0 ? wait_lo : wait_hi)) continue;
      while (ld_acquire_sys(flags + side) < epoch) {
         if (clock64() - t0 > timeout) { *err = 1 + side; return; }
         __nanosleep(200);
      }
   }
}

// copies the boundary planes of buffer S into the neighbours' halos (initial fill / after Monte Carlo colours)
__global__ void halo_push_kernel(int Nown, int M, size_t Npad, const SpinVec* __restrict__ S, EdgeParams ep) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (i < Nown) {
      const int lo = __ldg(ep.hdst_lo + i), hi = __ldg(ep.hdst_hi + i);
      if (lo >= 0 || hi >= 0) {
         const SpinVec v = S[(size_t)k * Npad + i];
         if (lo >= 0) ep.peer_lo[(size_t)k * Npad + lo] = v;
         if (hi >= 0) ep.peer_hi[(size_t)k * Npad + hi] = v;
      }
   }
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

// ------------------------------------------------------------------------------------------------
// The integrator of one site: midpoint (SOLVER 1) or Depondt (SOLVER 5), predictor (STAGE 1) or corrector +
// moment update (STAGE 2).  b = effective field at the spin `own` the field was evaluated with; c0 = spin at
// time t.  Returns the spin to store (pred in stage 1, cur in stage 2).
// ------------------------------------------------------------------------------------------------
// gpre: the three N(0,1) numbers of this (atom, ensemble, step) when the caller already drew them (asd_runs.cuh
// draws them while the gather list is in flight), else null.
// FROZEN (compile time): the kernel honours LlgParams::frozen.  Only the one-atom-per-thread direct kernel and the resident
// kernel are compiled with it (the check costs the register-blocked run kernel 1.5 %, measured); the engine routes
// fixed-moment runs to them.
// LEAN (compile time): the caller guarantees uniform damping / g factor / temperature (per_site == 0), no torque field and
// mompar == 0; the run kernel's plain-Heisenberg instantiation uses it (the runtime checks cost its serial part 30 %).
template <int SOLVER, int STAGE, bool FROZEN = false, bool LEAN = false>
__device__ __forceinline__ SpinVec integrate_site(const Tables& t, const LlgParams& p, int i, int k, int io, const double b[3],
                                                  const SpinVec& own, const SpinVec& c0, double* __restrict__ b2eff,
                                                  const float* gpre = nullptr, unsigned long long step_arg = ~0ull) {
   // step_arg: the noise key of a kernel that advances the step itself (llg_resident_kernel); default: p.step
   const unsigned long long nstep = (step_arg != ~0ull) ? step_arg : p.step;
   if (FROZEN && p.frozen != nullptr && __ldg(p.frozen + i)) {
      // not in red_atom_list: the loops of midpoint.f90:123 / depondt.f90:138 never visit this atom, emom2 keeps the value
      // magninit gave it (emom2 = emom, magnetizationinit.f90:538) and copym writes that back every step
      SpinVec o = c0;
      if (STAGE == 2 && p.mompar) o.m = calcm(p.mompar, c0.m, __ldg(p.mmom0 + (size_t)k * t.Npad + i), c0.z);
      return o;
   }
   const bool per_site = !LEAN && p.per_site;
   const int mompar = LEAN ? 0 : p.mompar;
   double lam, lg, temp;
   if (per_site) { lam = __ldg(p.lambda_a + i); lg = __ldg(p.landeg_a + i); temp = __ldg(p.temp_a + i); }
   else { lam = p.lambda; lg = p.landeg; temp = p.temp; }
   const double e[3] = {c0.x, c0.y, c0.z};
   const double m = c0.m;
   double g[3] = {0.0, 0.0, 0.0};
   if (gpre) { g[0] = (double)gpre[0]; g[1] = (double)gpre[1]; g[2] = (double)gpre[2]; }
   else if (p.thermal) gauss3f(p.seed, (uint32_t)io + t.atom_offset, (uint32_t)k + t.ens_offset, nstep, 0u, g[0], g[1], g[2]);
   double bt[3] = {0.0, 0.0, 0.0};
   if (!LEAN && t.btorque) {
      const double* __restrict__ q = t.btorque + (size_t)k * 3 * t.Npad + i;
      bt[0] = __ldg(q); bt[1] = __ldg(q + t.Npad); bt[2] = __ldg(q + 2 * (size_t)t.Npad);
   }
   const double lldamp = per_site ? 1.0 / (1.0 + lam * lam) : p.u_lldamp;
   SpinVec o;
   if (SOLVER == 1) {
      // ---- Mentink's semi-implicit midpoint (midpoint.f90) ----
      const double dt = per_site ? p.delta_t * 1.0 * p.gamma * lldamp : p.u_dt;  // bn = 1
      const double sqrtdt = per_site ? sqrt(dt) : p.u_sqrtdt;
      const double dtg = dt * lg, sqrtdtg = sqrtdt * lg;
      // rannum (randomnumbers.f90:667-670,735-746): sigma = sqrt(2 D), D = lam/(1+lam^2) k_B/(gamma mu_B) gamma / m * T
      double sigma = 0.0;
      if (p.thermal) {
         // sigma = sqrt(2 D), D = Dk / m * T * temprescale, evaluated as sqrt(2 Dk T temprescale) * m^-1/2
         double sg = p.u_sig1;
         if (per_site) sg = sqrt(2.0 * ((lam / (1 + lam * lam) * p.k_bolt / p.gamma / (p.mub)) * (p.gamma / 1.0)) * temp * p.temprescale);
         sigma = sg * rsqrt_nr(m);
      }
      const double r[3] = {g[0] * sigma, g[1] * sigma, g[2] * sigma};
      // etp: the spin the torque is evaluated with (old spin in stage 1, midpoint spin in stage 2)
      const double etp[3] = {own.x, own.y, own.z};
      double a1[3], s1[3], A[3];
      a1[0] = -bt[0] - b[0] - lam * (etp[1] * b[2] - etp[2] * b[1]);
      a1[1] = -bt[1] - b[1] - lam * (etp[2] * b[0] - etp[0] * b[2]);
      a1[2] = -bt[2] - b[2] - lam * (etp[0] * b[1] - etp[1] * b[0]);
      s1[0] = -r[0] - lam * (etp[1] * r[2] - etp[2] * r[1]);
      s1[1] = -r[1] - lam * (etp[2] * r[0] - etp[0] * r[2]);
      s1[2] = -r[2] - lam * (etp[0] * r[1] - etp[1] * r[0]);
#pragma unroll
      for (int a = 0; a < 3; a++) A[a] = 0.5 * dtg * a1[a] + 0.5 * sqrtdtg * s1[a];
      double et[3];
      cayley(e, A, et);
      if (STAGE == 1) {
         o.x = 0.5 * (e[0] + et[0]); o.y = 0.5 * (e[1] + et[1]); o.z = 0.5 * (e[2] + et[2]); o.m = m;
      } else {
         o.x = et[0]; o.y = et[1]; o.z = et[2];
         o.m = mompar ? calcm(mompar, m, __ldg(p.mmom0 + (size_t)k * t.Npad + i), et[2]) : m;
      }
   } else {
      // ---- Depondt (depondt.f90) ----
      double sigma = 0.0;
      if (p.thermal) {
         // sigma = sqrt(Dp temprescale T / m), evaluated as sqrt(Dp temprescale T) * m^-1/2
         double sg = p.u_sig5;
         if (per_site) sg = sqrt((2.0 * lam * p.k_bolt) / (p.delta_t * p.gamma * p.mub) * p.temprescale * temp);
         sigma = sg * rsqrt_nr(m);
      }
      const double bl[3] = {b[0] + g[0] * sigma, b[1] + g[1] * sigma, b[2] + g[2] * sigma};
      // damping cross product uses the spin the field was evaluated with (old spin / predictor spin)
      const double ep[3] = {own.x, own.y, own.z};
      double bd[3];
      bd[0] = bt[0] + bl[0] + lam * ep[1] * bl[2] - lam * ep[2] * bl[1];
      bd[1] = bt[1] + bl[1] + lam * ep[2] * bl[0] - lam * ep[0] * bl[2];
      bd[2] = bt[2] + bl[2] + lam * ep[0] * bl[1] - lam * ep[1] * bl[0];
      double* __restrict__ b2 = b2eff + (size_t)k * 3 * t.Npad + i;
      const double rot = p.delta_t * p.gamma * lldamp;
      double out[3];
      if (STAGE == 1) {
         rodrigues(bd, e, rot, out);
         b2[0] = bd[0]; b2[t.Npad] = bd[1]; b2[2 * (size_t)t.Npad] = bd[2];
         o.x = out[0]; o.y = out[1]; o.z = out[2]; o.m = m;
      } else {
         bd[0] = 0.5 * bd[0] + 0.5 * b2[0];
         bd[1] = 0.5 * bd[1] + 0.5 * b2[t.Npad];
         bd[2] = 0.5 * bd[2] + 0.5 * b2[2 * (size_t)t.Npad];
         rodrigues(bd, e, rot, out);
         o.x = out[0]; o.y = out[1]; o.z = out[2];
         o.m = mompar ? calcm(mompar, m, __ldg(p.mmom0 + (size_t)k * t.Npad + i), out[2]) : m;
      }
   }
   return o;
}

// ------------------------------------------------------------------------------------------------
// One stage of one LLG step, field evaluation fused with the integrator.
//   SOLVER 1 = semi-implicit midpoint, 5 = Depondt.   STAGE 1 = predictor, 2 = corrector (+ moment update).
//   STAGE 1: gathers from `cur`, writes `pred` (midpoint spin for SOLVER 1, rotated spin for SOLVER 5);
//   STAGE 2: gathers from `pred`, reads own `cur`, writes the new spin to `cur` (own slot only).
//   b2eff (Depondt only): [M][3][Npad] predictor field kept for the Heun average.
//   STAGED: tile path (shared-memory gather list).  EDGE: boundary tiles of a slab, see EdgeParams.
//   MSUM: the launch also leaves the per-tile sums of the new emomM in p.msum_part (corrector launches only).
// ------------------------------------------------------------------------------------------------
// ILEAN: the lean integrator (uniform LLG parameters, no torque field, mompar 0: checked by the engine per launch).
template <int SOLVER, int STAGE, bool REDUCED, bool STAGED, bool EDGE, bool MSUM, bool ILEAN = false>
__global__ void __launch_bounds__(256, STAGED ? ASD_MINB_STAGED : ASD_MINB)
llg_stage_kernel(const __grid_constant__ Tables t, const __grid_constant__ LlgParams p, const __grid_constant__ EdgeParams ep,
                 const TileRange tr, SpinVec* __restrict__ cur, SpinVec* __restrict__ pred, double* __restrict__ b2eff) {
   // programmatic dependent launch (asd_engine.cu, launch_dep): the next stage's CTAs may be scheduled once every CTA of this
   // grid has started; nothing below runs before the previous kernel of the stream has completed (no-ops otherwise)
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
   asm volatile("griddepcontrol.wait;" ::: "memory");
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   const int tile = ((int)blockIdx.x < tr.split) ? tr.first + (int)blockIdx.x : tr.second + ((int)blockIdx.x - tr.split);
   const int i = tile * 256 + threadIdx.x;
   const int k = blockIdx.y;
   SpinVec* __restrict__ curk = cur + (size_t)k * t.Npad;
   SpinVec* __restrict__ predk = pred + (size_t)k * t.Npad;
   const SpinVec* __restrict__ S = (STAGE == 1) ? curk : predk;
   // padding slots and the tail of the last tile take part in the staging but compute nothing
   int ih = 0, io = -1;
   bool active = i < t.Nown;
   double bs[3] = {0.0, 0.0, 0.0}, bq[3] = {0.0, 0.0, 0.0};
   double mnew[3] = {0.0, 0.0, 0.0};   // emomM written by this thread (fused observable)
   SpinVec own;
   uint4 w[ASD_NPF];
   double* __restrict__ s3 = nullptr;
   if (STAGED) {
      // (1) every independent global load in flight first: index words, {ham, orig}, own spins, gather list;
      // (2) stage emomM of the tile's gather list in shared memory; (3) sum from shared memory
      const int ii = active ? i : 0;
      idx_prologue(t, ii, t.zq8, w);       // words beyond this atom's list length hold the atom itself: harmless
      const int2 mt = __ldg(t.meta + ii);
      own = S[ii];
      if (STAGE == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(curk + ii));   // old spin, read after the sum
      const int cnt = __ldg(t.ucount + tile);
      const int* __restrict__ ul = t.ulist + (size_t)tile * t.ucap;
      prefetch_tile_staged(t, tile);
      const int ncpl = (t.cpl_param ? 0 : t.sm_cp) + t.sm_dm + t.sm_bq;
      s3 = sm + ncpl;
      for (int u0 = threadIdx.x; u0 < cnt; u0 += 3 * 256) {
         int sl[3];
#pragma unroll
         for (int a = 0; a < 3; a++) sl[a] = (u0 + a * 256 < cnt) ? __ldg(ul + u0 + a * 256) : 0;
         SpinVec v[3];
#pragma unroll
         for (int a = 0; a < 3; a++) v[a] = S[sl[a]];
#pragma unroll
         for (int a = 0; a < 3; a++)
            if (u0 + a * 256 < cnt) {
               double* __restrict__ m = s3 + 3 * (u0 + a * 256);
               m[0] = v[a].x * v[a].m; m[1] = v[a].y * v[a].m; m[2] = v[a].z * v[a].m;
            }
      }
      ih = REDUCED ? mt.x : 0;
      io = mt.y;
      active = active && io >= 0;
      stage_couplings(t, sm, smc, smd, smb);   // ends with __syncthreads() when it stages anything
      if (ncpl == 0) __syncthreads();
   } else {
      if (active) {
         io = __ldg(t.orig + i);
         ih = REDUCED ? __ldg(t.ham + i) : 0;
         active = io >= 0;
      }
      prefetch_tile(t, S, tile);
      stage_couplings(t, sm, smc, smd, smb);
   }
   if (active) {
      if (STAGED) {
         const int n = REDUCED ? __ldg(t.lsize + ih) : t.z;
         const int nq = (n + 7) >> 3;
         exchange_staged<REDUCED>(t, s3, i, ih, n, nq, w, smc, bs[0], bs[1], bs[2]);
         site_field<REDUCED, false>(t, S, i, ih, own, smc, smd, smb, bs, bq);
      } else {
         own = S[i];
         site_field<REDUCED>(t, S, i, ih, own, smc, smd, smb, bs, bq);
      }
      double h[3];
      ext_field(t, i, k, h);
      h[0] += p.tf[0]; h[1] += p.tf[1]; h[2] += p.tf[2];
      // beff = beff1 + beff2, beff2 = beff_q + external_field (hamiltonianactions.f90:240-243)
      const double b[3] = {bs[0] + (bq[0] + h[0]), bs[1] + (bq[1] + h[1]), bs[2] + (bq[2] + h[2])};
      SpinVec old;
      if (STAGE == 2) old = curk[i];
      const SpinVec o = integrate_site<SOLVER, STAGE, !STAGED, ILEAN>(t, p, i, k, io, b, own, (STAGE == 1) ? own : old, b2eff);
      if (STAGE == 1) predk[i] = o; else curk[i] = o;
      if (MSUM) { mnew[0] = o.x * o.m; mnew[1] = o.y * o.m; mnew[2] = o.z * o.m; }
      if (EDGE) {
         const int lo = __ldg(ep.hdst_lo + i), hi = __ldg(ep.hdst_hi + i);
         if (lo >= 0) ep.peer_lo[(size_t)k * t.Npad + lo] = o;
         if (hi >= 0) ep.peer_hi[(size_t)k * t.Npad + hi] = o;
      }
   }
   if (MSUM) {
      // per-tile sum of the new emomM (fixed tree: lanes -> warps -> CTA), prn_averages.f90:437-447
      __shared__ double red[3][8];
      const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
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
            for (int q = 0; q < 8; q++) v += red[threadIdx.x][q];
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

// ------------------------------------------------------------------------------------------------
// Resident kernel for SMALL systems (the reference's own regression cases: 43 ... a few thousand atoms).  Two launches
// per step cost several microseconds of launch latency and start with cold L1 caches, while the arithmetic of such a
// step is a short dependent chain per atom, so the whole time loop moves into ONE launch.  A thread-block CLUSTER of
// C <= 8 CTAs owns one ensemble; every CTA keeps cur / pred of ALL atoms of the ensemble in its own shared memory
// (2 x 32 bytes per slot) for the entire call and integrates the atoms i = rank * 256 + thread (+ C * 256 ...).  A new
// spin is stored into the copies of all C CTAs through distributed shared memory, so every gather is a LOCAL
// shared-memory read; predictor and corrector are separated by cluster barriers instead of kernel boundaries.  The
// neighbour table stays in global memory and is served from L1 after the first step.  Same device functions, same
// summation order (j = 1..nlistsize) and the same noise counters as llg_stage_kernel: a run is bit-identical whichever
// way it is launched.  t.nl4 / t.cpl_param / t.staged must be cleared by the caller.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_spin_cluster(cooperative_groups::cluster_group& cl, SpinVec* __restrict__ local, int i,
                                                   const SpinVec& v, unsigned nrank) {
   if (nrank == 1) { local[i] = v; return; }
   for (unsigned r = 0; r < nrank; r++) cl.map_shared_rank(local, r)[i] = v;
}

// cluster.sync() invalidates L1 (CCTL.IVALL): a single-CTA "cluster" takes the CTA barrier instead
__device__ __forceinline__ void resident_barrier(cooperative_groups::cluster_group& cl, unsigned nrank) {
   if (nrank == 1) __syncthreads(); else cl.sync();
}

// Pair sums of one atom with the neighbour slots read from the CTA's shared-memory copy of its atoms' lists
// (idx[(j) * nt], 16-bit slots): same terms, same order, same helpers as site_field.
template <bool REDUCED>
__device__ __forceinline__ void resident_pairs(const Tables& t, const SpinVec* __restrict__ S, const unsigned short* __restrict__ idx,
                                               int nt, int i, int ih, const SpinVec& own, const double* smc, const double* smd,
                                               const double* smb, double bs[3], double bq[3]) {
   const int Npad = t.Npad;
   double fx = 0.0, fy = 0.0, fz = 0.0;
   {
      const int n = REDUCED ? __ldg(t.lsize + ih) : t.z;
      const double* __restrict__ c = REDUCED ? (smc ? smc + (size_t)ih * t.z : t.cp + (size_t)ih * t.z) : t.cp + i;
#pragma unroll 10
      for (int j = 0; j < n; j++) {
         const SpinVec v = S[idx[j * nt]];
         const double cj = REDUCED ? c[j] : __ldg(c + (size_t)j * Npad);
         fx = fma(cj, v.x * v.m, fx);
         fy = fma(cj, v.y * v.m, fy);
         fz = fma(cj, v.z * v.m, fz);
      }
   }
   idx += t.z * nt;
   if (t.zdm > 0) {
      const int n = REDUCED ? __ldg(t.dmsize + ih) : t.zdm;
      for (int j = 0; j < n; j++) {
         double Dx, Dy, Dz;
         if (REDUCED) {
            const double* __restrict__ d = (smd ? smd : t.dmv) + ((size_t)ih * t.zdm + j) * 3;
            Dx = d[0]; Dy = d[1]; Dz = d[2];
         } else {
            const size_t o = (size_t)j * Npad + i, s = (size_t)t.zdm * Npad;
            Dx = __ldg(t.dmv + o); Dy = __ldg(t.dmv + s + o); Dz = __ldg(t.dmv + 2 * s + o);
         }
         const SpinVec v = S[idx[j * nt]];
         dm_term(Dx, Dy, Dz, v.x * v.m, v.y * v.m, v.z * v.m, fx, fy, fz);
      }
   }
   idx += t.zdm * nt;
   double qx = 0.0, qy = 0.0, qz = 0.0;
   if (t.zbq > 0) {
      const double ox = own.x * own.m, oy = own.y * own.m, oz = own.z * own.m;
      const int n = REDUCED ? __ldg(t.bqsize + ih) : t.zbq;
      for (int j = 0; j < n; j++) {
         const double jb = REDUCED ? (smb ? smb : t.jbq)[(size_t)ih * t.zbq + j] : __ldg(t.jbq + (size_t)j * Npad + i);
         const SpinVec v = S[idx[j * nt]];
         bq_term(jb, v.x * v.m, v.y * v.m, v.z * v.m, ox, oy, oz, qx, qy, qz);
      }
   }
   bs[0] = fx; bs[1] = fy; bs[2] = fz;
   bq[0] = qx; bq[1] = qy; bq[2] = qz;
}

// apt = atoms per thread (host: ceil(Nown / (nrank * blockDim.x))); shared memory = staged couplings | cur | pred |
// 16-bit neighbour slots [apt][z + zdm + zbq][blockDim.x] of the atoms this CTA integrates
template <int SOLVER, bool REDUCED>
__global__ void __launch_bounds__(256, 1)
llg_resident_kernel(const __grid_constant__ Tables t, const __grid_constant__ LlgParams p, SpinVec* __restrict__ cur,
                    SpinVec* __restrict__ pred, double* __restrict__ b2eff, long nsteps, unsigned long long first_step, int apt) {
   namespace cg = cooperative_groups;
   cg::cluster_group cl = cg::this_cluster();
   const unsigned nrank = cl.num_blocks(), rank = cl.block_rank();
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   stage_couplings(t, sm, smc, smd, smb);
   const int ncpl = t.sm_cp + t.sm_dm + t.sm_bq;
   SpinVec* __restrict__ shc = reinterpret_cast<SpinVec*>(sm + ((ncpl + 3) & ~3));   // 32-byte aligned
   SpinVec* __restrict__ shp = shc + t.Npad;
   unsigned short* __restrict__ idx = reinterpret_cast<unsigned short*>(shp + t.Npad);
   const int nt = blockDim.x, zt = t.z + t.zdm + t.zbq;
   const int k = blockIdx.x / nrank;
   SpinVec* __restrict__ curk = cur + (size_t)k * t.Npad;
   for (int i = threadIdx.x; i < t.Npad; i += nt) { const SpinVec v = curk[i]; shc[i] = v; shp[i] = v; }
   const int first = rank * nt + threadIdx.x, stride = nrank * nt;
   for (int a = 0; a < apt; a++) {
      const int i = first + a * stride;
      unsigned short* __restrict__ row = idx + (size_t)a * zt * nt + threadIdx.x;
      if (i < t.Nown) {
         for (int j = 0; j < t.z; j++) row[j * nt] = (unsigned short)__ldg(t.nl + (size_t)j * t.Npad + i);
         for (int j = 0; j < t.zdm; j++) row[(t.z + j) * nt] = (unsigned short)__ldg(t.dml + (size_t)j * t.Npad + i);
         for (int j = 0; j < t.zbq; j++) row[(t.z + t.zdm + j) * nt] = (unsigned short)__ldg(t.bql + (size_t)j * t.Npad + i);
      }
   }
   resident_barrier(cl, nrank);
   for (long s = 0; s < nsteps; s++) {
      const unsigned long long step = first_step + (unsigned long long)s;
      // ---- field(cur) + predictor ----
      for (int a = 0; a < apt; a++) {
         const int i = first + a * stride;
         if (i >= t.Nown) break;
         const int io = __ldg(t.orig + i);
         if (io < 0) continue;
         const int ih = REDUCED ? __ldg(t.ham + i) : 0;
         const SpinVec own = shc[i];
         double bs[3], bq[3], h[3];
         resident_pairs<REDUCED>(t, shc, idx + (size_t)a * zt * nt + threadIdx.x, nt, i, ih, own, smc, smd, smb, bs, bq);
         site_field<REDUCED, false, ASD_CHUNK, false, false>(t, shc, i, ih, own, smc, smd, smb, bs, bq);
         ext_field(t, i, k, h);
         add_time_field(p, step, h);
         const double b[3] = {bs[0] + (bq[0] + h[0]), bs[1] + (bq[1] + h[1]), bs[2] + (bq[2] + h[2])};
         store_spin_cluster(cl, shp, i, integrate_site<SOLVER, 1, true>(t, p, i, k, io, b, own, own, b2eff, nullptr, step), nrank);
      }
      resident_barrier(cl, nrank);
      // ---- field(pred) + corrector + moment update: cur[i] is read by its owner only during this stage ----
      for (int a = 0; a < apt; a++) {
         const int i = first + a * stride;
         if (i >= t.Nown) break;
         const int io = __ldg(t.orig + i);
         if (io < 0) continue;
         const int ih = REDUCED ? __ldg(t.ham + i) : 0;
         const SpinVec own = shp[i], old = shc[i];
         double bs[3], bq[3], h[3];
         resident_pairs<REDUCED>(t, shp, idx + (size_t)a * zt * nt + threadIdx.x, nt, i, ih, own, smc, smd, smb, bs, bq);
         site_field<REDUCED, false, ASD_CHUNK, false, false>(t, shp, i, ih, own, smc, smd, smb, bs, bq);
         ext_field(t, i, k, h);
         add_time_field(p, step, h);
         const double b[3] = {bs[0] + (bq[0] + h[0]), bs[1] + (bq[1] + h[1]), bs[2] + (bq[2] + h[2])};
         store_spin_cluster(cl, shc, i, integrate_site<SOLVER, 2, true>(t, p, i, k, io, b, own, old, b2eff, nullptr, step), nrank);
      }
      resident_barrier(cl, nrank);
   }
   if (rank == 0) {
      SpinVec* __restrict__ predk = pred + (size_t)k * t.Npad;
      for (int i = threadIdx.x; i < t.Npad; i += nt) { curk[i] = shc[i]; predk[i] = shp[i]; }
   }
}

// ------------------------------------------------------------------------------------------------
// Field-only kernel (effective_field_full): writes beff / beff1 / beff2 in ORIGINAL atom order,
// Fortran shape (3,N,M), and the per-site energy term (hamiltonianactions.f90:245-250) for reduction.
// ------------------------------------------------------------------------------------------------
template <bool REDUCED>
__global__ void __launch_bounds__(256)
field_kernel(const __grid_constant__ Tables t, const SpinVec* __restrict__ cur, double* __restrict__ beff, double* __restrict__ beff1,
             double* __restrict__ beff2, double* __restrict__ esite) {
   extern __shared__ double sm[];
   const double *smc, *smd, *smb;
   stage_couplings(t, sm, smc, smd, smb);
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const int k = blockIdx.y;
   if (i >= t.Npad) return;
   const int o = __ldg(t.orig + i);
   if (o < 0) return;
   const int ih = REDUCED ? __ldg(t.ham + i) : 0;
   const SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
   const SpinVec own = S[i];
   double bs[3], bq[3], h[3];
   site_field<REDUCED>(t, S, i, ih, own, smc, smd, smb, bs, bq);
   ext_field(t, i, k, h);
   const size_t q = 3 * ((size_t)o + (size_t)t.N * k);
#pragma unroll
   for (int a = 0; a < 3; a++) {
      const double b2 = bq[a] + h[a];
      if (beff1) beff1[q + a] = bs[a];
      if (beff2) beff2[q + a] = b2;
      if (beff) beff[q + a] = bs[a] + b2;
   }
   if (esite) {
      const double mx = own.x * own.m, my = own.y * own.m, mz = own.z * own.m;
      const double tx = 0.5 * (bs[0] + 2.0 * bq[0] + 2.0 * h[0]);
      const double ty = 0.5 * (bs[1] + 2.0 * bq[1] + 2.0 * h[1]);
      const double tz = 0.5 * (bs[2] + 2.0 * bq[2] + 2.0 * h[2]);
      esite[(size_t)k * t.Npad + i] = -mx * tx - my * ty - mz * tz;
   }
}

// ------------------------------------------------------------------------------------------------
// Observables: per-ensemble sum of emomM (and of esite) with a fixed-shape two-pass tree reduction
// (deterministic for a given Npad): pass 1 = one partial per CTA, pass 2 = one CTA per ensemble.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
   return v;
}

__global__ void __launch_bounds__(256)
moment_partial_kernel(int Npad, const int* __restrict__ orig, const SpinVec* __restrict__ cur,
                      const double* __restrict__ esite, double* __restrict__ part /*[M][gridDim.x][4]*/) {
   __shared__ double red[4][8];
   const int k = blockIdx.y;
   double s[4] = {0, 0, 0, 0};
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) {
      if (__ldg(orig + i) < 0) continue;
      const SpinVec v = cur[(size_t)k * Npad + i];
      s[0] += v.x * v.m; s[1] += v.y * v.m; s[2] += v.z * v.m;
      if (esite) s[3] += esite[(size_t)k * Npad + i];
   }
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
   for (int a = 0; a < 4; a++) { s[a] = warp_sum(s[a]); if (l == 0) red[a][w] = s[a]; }
   __syncthreads();
   if (w == 0) {
#pragma unroll
      for (int a = 0; a < 4; a++) {
         double v = (l < 8) ? red[a][l] : 0.0;
         v = warp_sum(v);
         if (l == 0) part[((size_t)k * gridDim.x + blockIdx.x) * 4 + a] = v;
      }
   }
}

__global__ void __launch_bounds__(1024)
moment_final_kernel(int nblk, const double* __restrict__ part, double* __restrict__ out /*[M][4]*/) {
   // one CTA per ensemble; every thread keeps several independent 32-byte loads in flight (the partials of up to a
   // few 10^4 tiles are summed in a fixed order: thread-strided, then lanes -> warps -> CTA)
   __shared__ double red[4][32];
   const int k = blockIdx.x;
   const double4* __restrict__ src = reinterpret_cast<const double4*>(part) + (size_t)k * nblk;
   double s[4] = {0, 0, 0, 0};
#pragma unroll 8
   for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
      const double4 v = src[b];
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
   }
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
   for (int a = 0; a < 4; a++) { s[a] = warp_sum(s[a]); if (l == 0) red[a][w] = s[a]; }
   __syncthreads();
   if (w == 0) {
      const int nw = blockDim.x >> 5;
#pragma unroll
      for (int a = 0; a < 4; a++) {
         double v = (l < nw) ? red[a][l] : 0.0;
         v = warp_sum(v);
         if (l == 0) out[k * 4 + a] = v;
      }
   }
}

// ------------------------------------------------------------------------------------------------
// Term-resolved energy (calc_energy, source/Hamiltonian/energy.f90:180-340): per site
//   exc = -1/2 m.B_xc, dm = -1/2 m.B_dm, bq = -1/4 m.B_bq, ani = -1/2 m.B_ani, ext = -m.B_ext   (update_ene factors)
// written as rows [k][5][Npad] (order: exc, ani, dm, bq, ext = the columns of totenergy.*.out that this path has);
// reduced per ensemble by reduce_rows_*.  Measurement-only kernel: plain gathers.
// ------------------------------------------------------------------------------------------------
template <bool REDUCED>
__global__ void __launch_bounds__(256)
energy_terms_kernel(const __grid_constant__ Tables t, const SpinVec* __restrict__ cur, double* __restrict__ rows) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const int k = blockIdx.y;
   if (i >= t.Npad) return;
   double* __restrict__ out = rows + (size_t)k * 5 * t.Npad + i;
   const int o = __ldg(t.orig + i);
   if (o < 0) { for (int a = 0; a < 5; a++) out[(size_t)a * t.Npad] = 0.0; return; }
   const int ih = REDUCED ? __ldg(t.ham + i) : 0;
   const int Npad = t.Npad;
   const SpinVec* __restrict__ S = cur + (size_t)k * t.Npad;
   const SpinVec own = S[i];
   const double ox = own.x * own.m, oy = own.y * own.m, oz = own.z * own.m;
   double f[3] = {0.0, 0.0, 0.0};
   {  // Heisenberg
      const int n = REDUCED ? __ldg(t.lsize + ih) : t.z;
      for (int j = 0; j < n; j++) {
         const SpinVec v = S[__ldg(t.nl + (size_t)j * Npad + i)];
         if (t.jtens) {
            const double mx = v.x * v.m, my = v.y * v.m, mz = v.z * v.m;
            double J[9];
            for (int a = 0; a < 9; a++)
               J[a] = REDUCED ? __ldg(t.cp + ((size_t)ih * t.z + j) * 9 + a) : __ldg(t.cp + ((size_t)a * t.z + j) * Npad + i);
            f[0] = f[0] + J[0] * mx + J[3] * my + J[6] * mz;
            f[1] = f[1] + J[1] * mx + J[4] * my + J[7] * mz;
            f[2] = f[2] + J[2] * mx + J[5] * my + J[8] * mz;
            continue;
         }
         const double cj = REDUCED ? __ldg(t.cp + (size_t)ih * t.z + j) : __ldg(t.cp + (size_t)j * Npad + i);
         f[0] = fma(cj, v.x * v.m, f[0]); f[1] = fma(cj, v.y * v.m, f[1]); f[2] = fma(cj, v.z * v.m, f[2]);
      }
   }
   out[0] = -0.5 * (ox * f[0] + oy * f[1] + oz * f[2]);
   f[0] = f[1] = f[2] = 0.0;
   if (t.zdm > 0) {
      const int n = REDUCED ? __ldg(t.dmsize + ih) : t.zdm;
      for (int j = 0; j < n; j++) {
         const SpinVec v = S[__ldg(t.dml + (size_t)j * Npad + i)];
         double Dx, Dy, Dz;
         if (REDUCED) { const double* __restrict__ d = t.dmv + ((size_t)ih * t.zdm + j) * 3; Dx = d[0]; Dy = d[1]; Dz = d[2]; }
         else { const size_t q = (size_t)j * Npad + i, s = (size_t)t.zdm * Npad; Dx = __ldg(t.dmv + q); Dy = __ldg(t.dmv + s + q); Dz = __ldg(t.dmv + 2 * s + q); }
         const double mx = v.x * v.m, my = v.y * v.m, mz = v.z * v.m;
         f[0] = f[0] + Dz * my - Dy * mz; f[1] = f[1] + Dx * mz - Dz * mx; f[2] = f[2] + Dy * mx - Dx * my;
      }
   }
   out[(size_t)2 * Npad] = -0.5 * (ox * f[0] + oy * f[1] + oz * f[2]);
   f[0] = f[1] = f[2] = 0.0;
   if (t.zbq > 0) {
      const int n = REDUCED ? __ldg(t.bqsize + ih) : t.zbq;
      for (int j = 0; j < n; j++) {
         const SpinVec v = S[__ldg(t.bql + (size_t)j * Npad + i)];
         const double jb = REDUCED ? __ldg(t.jbq + (size_t)ih * t.zbq + j) : __ldg(t.jbq + (size_t)j * Npad + i);
         const double mx = v.x * v.m, my = v.y * v.m, mz = v.z * v.m;
         const double c = 2.0 * jb * (mx * ox + my * oy + mz * oz);
         f[0] = fma(c, mx, f[0]); f[1] = fma(c, my, f[1]); f[2] = fma(c, mz, f[2]);
      }
   }
   out[(size_t)3 * Npad] = -0.25 * (ox * f[0] + oy * f[1] + oz * f[2]);
   f[0] = f[1] = f[2] = 0.0;
   if (t.do_aniso) {
      const int ta = __ldg(t.taniso + i);
      if (ta == 1 || ta == 2 || ta == 7) {
         const double k1 = __ldg(t.kaniso + i), k2 = __ldg(t.kaniso + Npad + i);
         if (ta == 1 || ta == 7) {
            const double ex = __ldg(t.eaniso + i), ey = __ldg(t.eaniso + Npad + i), ez = __ldg(t.eaniso + 2 * (size_t)Npad + i);
            const double tt1 = ox * ex + oy * ey + oz * ez;
            const double tt3 = 2.0 * tt1 * (k1 + 2.0 * k2 * (1.0 - tt1 * tt1));
            f[0] -= tt3 * ex; f[1] -= tt3 * ey; f[2] -= tt3 * ez;
         }
         if (ta == 2 || ta == 7) {
            const double x2 = ox * ox, y2 = oy * oy, z2 = oz * oz;
            const double s = (ta == 7) ? __ldg(t.sb + i) : 1.0;
            f[0] += s * (2.0 * k1 * ox * (y2 + z2) + 2.0 * k2 * ox * (y2 * z2));
            f[1] += s * (2.0 * k1 * oy * (z2 + x2) + 2.0 * k2 * oy * (z2 * x2));
            f[2] += s * (2.0 * k1 * oz * (x2 + y2) + 2.0 * k2 * oz * (x2 * y2));
         }
      }
   }
   out[(size_t)1 * Npad] = -0.5 * (ox * f[0] + oy * f[1] + oz * f[2]);
   double h[3];
   ext_field(t, i, k, h);
   out[(size_t)4 * Npad] = -(ox * h[0] + oy * h[1] + oz * h[2]);
}

// per-ensemble sums of R rows of length Npad: rows[k][R][Npad] -> part[k][R][gridDim.x] -> out[k][R] (fixed tree)
__global__ void __launch_bounds__(256)
reduce_rows_partial_kernel(int Npad, int R, const double* __restrict__ rows, double* __restrict__ part) {
   __shared__ double red[8];
   const int r = blockIdx.y, k = blockIdx.z;
   const double* __restrict__ src = rows + ((size_t)k * R + r) * Npad;
   double s = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) s += src[i];
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   s = warp_sum(s);
   if (l == 0) red[w] = s;
   __syncthreads();
   if (w == 0) {
      double v = (l < 8) ? red[l] : 0.0;
      v = warp_sum(v);
      if (l == 0) part[((size_t)k * R + r) * gridDim.x + blockIdx.x] = v;
   }
}

__global__ void __launch_bounds__(256)
reduce_rows_final_kernel(int nblk, const double* __restrict__ part, double* __restrict__ out) {
   __shared__ double red[8];
   const double* __restrict__ src = part + (size_t)blockIdx.x * nblk;
   double s = 0.0;
   for (int b = threadIdx.x; b < nblk; b += blockDim.x) s += src[b];
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
   s = warp_sum(s);
   if (l == 0) red[w] = s;
   __syncthreads();
   if (w == 0) {
      double v = (l < 8) ? red[l] : 0.0;
      v = warp_sum(v);
      if (l == 0) out[blockIdx.x] = v;
   }
}

// selected atoms (trajectory measurements, prn_trajectories.f90): out[k][n][4] = {ex, ey, ez, m} of atom slots sl[n]
__global__ void gather_atoms_kernel(int n, int M, size_t Npad, const int* __restrict__ sl, const SpinVec* __restrict__ cur,
                                    double* __restrict__ out) {
   const int q = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (q >= n) return;
   const SpinVec v = cur[(size_t)k * Npad + sl[q]];
   double* o = out + ((size_t)k * n + q) * 4;
   o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.m;
}

// Sublattice-projected moment sums (buffer_proj_avrg, prn_averages.f90:462-512): the partial sums of moment_partial_kernel
// restricted to the atoms whose basis number mod(i-1, NA) equals cls; same fixed-shape tree, same final kernel.
__global__ void __launch_bounds__(256)
moment_class_partial_kernel(int Npad, const int* __restrict__ orig, const SpinVec* __restrict__ cur, int NA, int cls,
                            double* __restrict__ part /*[M][gridDim.x][4]*/) {
   __shared__ double red[4][8];
   const int k = blockIdx.y;
   double s[4] = {0, 0, 0, 0};
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) {
      const int o = __ldg(orig + i);
      if (o < 0 || o % NA != cls) continue;
      const SpinVec v = cur[(size_t)k * Npad + i];
      s[0] += v.x * v.m; s[1] += v.y * v.m; s[2] += v.z * v.m;
   }
   const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
   for (int a = 0; a < 4; a++) { s[a] = warp_sum(s[a]); if (l == 0) red[a][w] = s[a]; }
   __syncthreads();
   if (w == 0) {
#pragma unroll
      for (int a = 0; a < 4; a++) {
         double v = (l < 8) ? red[a][l] : 0.0;
         v = warp_sum(v);
         if (l == 0) part[((size_t)k * gridDim.x + blockIdx.x) * 4 + a] = v;
      }
   }
}

// Solid angle of every triangle of a triangulated lattice (pontryagin_tri, source/Measurement/topology.f90:78-116):
// q = 2 atan( m1 . (m2 x m3) / (1 + m1.m2 + m1.m3 + m2.m3) ) with the UNIT moments of the three corners;
// rows[k][t] = q of triangle t in ensemble k (summed by the reduce_rows kernels).  tri[c][t] = slot of corner c.
__global__ void __launch_bounds__(256)
skyrmion_tri_kernel(int nsimp, size_t Npad, const int* __restrict__ tri, const SpinVec* __restrict__ cur, double* __restrict__ rows) {
   const int t = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (t >= nsimp) return;
   const SpinVec* __restrict__ S = cur + (size_t)k * Npad;
   const SpinVec a = S[__ldg(tri + t)], b = S[__ldg(tri + nsimp + t)], c = S[__ldg(tri + 2 * (size_t)nsimp + t)];
   // f_volume(m1, m2, m3) = (m1 x m2) . m3 (math_functions.f90:59-73)
   const double cx = a.y * b.z - a.z * b.y, cy = a.z * b.x - a.x * b.z, cz = a.x * b.y - a.y * b.x;
   const double vol = c.x * cx + c.y * cy + c.z * cz;
   const double ab = a.x * b.x + a.y * b.y + a.z * b.z;
   const double ac = a.x * c.x + a.y * c.y + a.z * c.z;
   const double bc = b.x * c.x + b.y * c.y + b.z * c.z;
   const double qq = vol / (1.0 + ab + ac + bc);
   rows[(size_t)k * nsimp + t] = 2.0 * atan(qq);
}

// ------------------------------------------------------------------------------------------------
// Layout conversion between the host's Fortran arrays and the packed device order.
// ------------------------------------------------------------------------------------------------
__global__ void pack_kernel(int N, int Nown, int Npad, int M, const int* __restrict__ orig, const double* __restrict__ emom,
                            const double* __restrict__ mmom, SpinVec* __restrict__ cur) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (i >= Nown) return;
   const int o = orig[i];
   SpinVec v;
   if (o < 0) { v.x = 0; v.y = 0; v.z = 1.0; v.m = 0.0; }
   else {
      const size_t q = (size_t)o + (size_t)N * k;
      v.x = emom[3 * q]; v.y = emom[3 * q + 1]; v.z = emom[3 * q + 2]; v.m = mmom[q];
   }
   cur[(size_t)k * Npad + i] = v;
}

__global__ void unpack_kernel(int N, int Npad, int M, const int* __restrict__ orig, const SpinVec* __restrict__ cur,
                              double* __restrict__ emom, double* __restrict__ emomM, double* __restrict__ mmom) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (i >= Npad) return;
   const int o = orig[i];
   if (o < 0) return;
   const SpinVec v = cur[(size_t)k * Npad + i];
   const size_t q = (size_t)o + (size_t)N * k;
   if (emom) { emom[3 * q] = v.x; emom[3 * q + 1] = v.y; emom[3 * q + 2] = v.z; }
   if (emomM) { emomM[3 * q] = v.x * v.m; emomM[3 * q + 1] = v.y * v.m; emomM[3 * q + 2] = v.z * v.m; }
   if (mmom) mmom[q] = v.m;
}

__global__ void zip_meta_kernel(int Npad, const int* __restrict__ ham, const int* __restrict__ orig, int2* __restrict__ meta) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s < Npad) meta[s] = make_int2(ham[s], orig[s]);
}

// vectorised copies of a slot-major table: nl[z][Npad] -> nl4[zq][Npad], cp[z][Npad] -> cp4[zq][Npad]
__global__ void vectorise_table_kernel(int Npad, int z, int zq, const int* __restrict__ nl, const double* __restrict__ cp,
                                       int4* __restrict__ nl4, double4* __restrict__ cp4) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
   if (s >= Npad) return;
   int v[4];
   double c[4];
   for (int a = 0; a < 4; a++) {
      const int j = 4 * q + a;
      v[a] = (j < z) ? nl[(size_t)j * Npad + s] : s;
      c[a] = (cp && j < z) ? cp[(size_t)j * Npad + s] : 0.0;
   }
   if (nl4) nl4[(size_t)q * Npad + s] = make_int4(v[0], v[1], v[2], v[3]);
   if (cp4) cp4[(size_t)q * Npad + s] = make_double4(c[0], c[1], c[2], c[3]);
}

// scatter a per-atom host-order array (ncomp,N[,M]) into device order [M][ncomp][Npad]
__global__ void scatter_kernel(int N, int Npad, int ncomp, int per_ens, const int* __restrict__ orig,
                               const double* __restrict__ src, double* __restrict__ dst) {
   const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (i >= Npad) return;
   const int o = orig[i];
   for (int a = 0; a < ncomp; a++) {
      double v = 0.0;
      if (o >= 0) v = src[a + (size_t)ncomp * ((size_t)o + (per_ens ? (size_t)N * k : 0))];
      dst[((size_t)k * ncomp + a) * Npad + i] = v;
   }
}

}  // namespace asd
