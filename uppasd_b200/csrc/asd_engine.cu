// Host side of libuppasd_b200.so: the engine that owns device memory, lays the reference's tables out for
// the B200 and sequences the fused kernels; plus the C ABI (include/uppasd_b200.h).
//
// Mirrors, on the host, the call order of the reference drivers for this path:
//   sd_mphase loop   source/sd_driver.f90:517-849  (measure -> field -> evolve_first -> field -> evolve_second
//                                                  -> moment_update)
//   native boundary  source/gpu_files/cudaMdSimulation.cu:300-512, fortranData.cpp:141-185, fort_helper.cpp
//   mc_mphase loop   source/mc_driver.f90:310-430  (measure -> mc_evolve)
// None of the reference's native code is reused: two fused kernels per LLG step instead of ~14 launches, a
// packed 32-byte spin layout in device order, counter-based in-register noise, colour-parallel MC.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/uppasd_b200.h"
#include "asd_device.cuh"
#include "asd_mc.cuh"
#include "asd_mc_block.cuh"
#include "asd_mc_runs.cuh"
#include "asd_tiles.cuh"
#include "asd_runs.cuh"
#include "asd_lattice.cuh"

using namespace asd;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
   char buf[1024];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_err = buf;
   return code;
}
#define CU(x)                                                                                          \
   do {                                                                                                \
      cudaError_t err__ = (x);                                                                         \
      if (err__ != cudaSuccess) return fail(-100, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(err__), \
                                            __FILE__, __LINE__, cudaGetErrorString(err__));            \
   } while (0)

template <class T>
struct DevBuf {
   T* p = nullptr;
   size_t n = 0;
   int alloc(size_t count) {
      if (p && n == count) return 0;   // keep the allocation (peers of a slab hold IPC mappings of cur / pred)
      release();
      if (count == 0) return 0;
      cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
      if (e != cudaSuccess) { p = nullptr; return fail(-101, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e)); }
      n = count;
      return 0;
   }
   int upload(const std::vector<T>& h, cudaStream_t s) {
      int r = alloc(h.size());
      if (r) return r;
      if (h.empty()) return 0;
      CU(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
      CU(cudaStreamSynchronize(s));
      return 0;
   }
   void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
   ~DevBuf() { release(); }
};

// host-side description of one pair-interaction table exactly as the reference holds it
struct HostTable {
   int z = 0, ncomp = 1;
   std::vector<int> list;      // (z, N) column-major, 1-based, 0 = empty
   std::vector<int> lsize;     // (NH)
   std::vector<double> coup;   // (ncomp, z, NH)
   bool present() const { return z > 0; }
};

// ------------------------------------------------------------------------------------------------
// Layout: one device ordering of the atoms + the tables permuted/transposed into it.
// ------------------------------------------------------------------------------------------------
struct Layout {
   DevBuf<unsigned char> d_frozen;   // [Npad] device order, only for fixed-moment runs
   int N = 0, Npad = 0, NH = 0, M = 0;
   bool reduced = false;
   std::vector<int> orig;     // [Npad] original 0-based atom of slot, -1 padding
   std::vector<int> slot_of;  // [N] slot of original atom
   // colour classes (MC layout only): slot ranges
   std::vector<int> colour_first, colour_count;
   DevBuf<int> d_ham, d_orig, d_nl, d_lsize, d_dml, d_dmsize, d_bql, d_bqsize, d_taniso;
   DevBuf<int4> d_nl4;
   DevBuf<int> d_nlrow;             // atom-major exchange table of the MC layout (mc_colour_coop_kernel)
   DevBuf<double> d_cprow;          // atom-major per-atom couplings next to it (non-reduced layouts)
   DevBuf<int2> d_classes;          // {first, count} of every colour class (mc_sweeps_persistent_kernel)
   DevBuf<int> d_ucount, d_ulist;   // staged tile path (asd_tiles.cuh)
   DevBuf<uint4> d_nl16, d_dm16, d_bq16;
   DevBuf<int2> d_meta;
   DevBuf<uint4> d_utab;            // run-compressed table (asd_runs.cuh)
   DevBuf<int> d_gcount;
   DevBuf<int> d_okey;              // sort key of the gather lists when it differs from orig (lattice builder)
   bool is_mc = false;
   DevBuf<double4> d_cp4;
   DevBuf<int> d_cnt[3];  // per-atom list lengths of device-built tables (exchange, DM, BQ)
   int zs[3] = {0, 0, 0};
   DevBuf<double> d_cp, d_dmv, d_jbq, d_eaniso, d_kaniso, d_sb, d_ext, d_btorque, d_landeg, d_lambda, d_temp, d_mmom0;
   Tables t{};
   size_t smem_bytes = 0;
};

// Slab decomposition of one supercell along z, one slab per engine / GPU.  The ring neighbours' cur,
// pred and flag buffers are mapped into this process (CUDA IPC, or plain pointers inside one process); the EDGE
// launches of the stage kernels store boundary spins straight into them.
struct Slab {
   int on = 0, G = 1, g = 0, H = 0;
   DevBuf<int> hdst_lo, hdst_hi;
   DevBuf<unsigned long long> flags;   // [0] published by the lower neighbour, [1] by the upper neighbour
   DevBuf<unsigned int> ctr;
   DevBuf<int> err;
   SpinVec* peer_cur[2] = {nullptr, nullptr};    // [0] lower neighbour, [1] upper neighbour
   SpinVec* peer_pred[2] = {nullptr, nullptr};
   double* peer_mcur[2] = {nullptr, nullptr};    // the neighbours' moment planes (MM run kernels push emomM next to the spin)
   double* peer_mpred[2] = {nullptr, nullptr};
   unsigned long long* peer_flags[2] = {nullptr, nullptr};
   void* opened[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
   int n_opened = 0;
   unsigned long long epoch = 0;       // exchanges completed (identical on every rank)
   bool connected = false;
   // moment-plane layouts: the halo exchange of a stage is its own small launch on a high-priority side stream, concurrent with the
   // interior tiles of the stage (evB: boundary tiles done -> push may start; evP: push done -> its source buffer may be rewritten)
   cudaStream_t push_stream = nullptr;
   cudaEvent_t evB = nullptr, evP = nullptr;
   bool push_pending = false;
   long long timeout_ticks = 8000000000LL;  // ~4 s at 1.9 GHz
};

// unit-cell stencil of a device-built table (asd_build_lattice_table), kept for the periodic colouring
struct Stencil {
   int maxslot = 0;
   std::vector<int> nslot, cell_atom, cell_shift;
   bool present() const { return maxslot > 0; }
};

// tables of the Monte Carlo block sweep (asd_mc_block.cuh), built lazily on the SD (brick) layout of a device-built lattice
struct McBlockState {
   bool tried = false, on = false;
   int ts = 0, ucap = 0, ncol = 0, ntile = 0;
   size_t smem = 0;
   std::vector<int> class_first, class_count, h_tilelist;   // tile-colour classes: ranges of h_tilelist
   std::vector<int> class_run;                              // leading tiles of every class that take the run form (asd_mc_runs.cuh)
   McRuns mr{};
   size_t smem_run = 0;
   int nt = 256;                                            // threads per CTA of the run form
   bool ticket = false;                                     // one launch per sweep with tile dependencies (McTicket)
   int adjcap = 0;
   unsigned long long tickets = 0;
   unsigned int epoch = 0;
   DevBuf<int> adj, nadj;
   DevBuf<unsigned char> tclass;
   DevBuf<unsigned int> done;
   DevBuf<unsigned long long> counter;
   DevBuf<unsigned short> gtab;
   DevBuf<int> ulist, ucount, cstart, tilelist;
   DevBuf<uint4> nl16, dm16, bq16;
   DevBuf<unsigned short> selfpos, corder;
   DevBuf<double> drec;                                     // [M][ntile][1024][5] pre-drawn trial moves of the running sweep
};

struct asd_engine {
   int device = 0;
   McBlockState mcb;
   Slab slab;
   cudaStream_t stream = nullptr;
   // constants
   double gamma = 1.760859644e11, k_bolt = 1.38064852e-23, mub = 9.274009994e-24, mry = 2.179872325e-21;
   // system
   int N = 0, M = 0, NH = 0;
   std::vector<int> aHam;  // 1-based ham row per atom
   HostTable ex, dm, bq;
   bool jtensor = false;   // ex holds j_tens(3,3,z,NH): tensorial exchange (do_jtensor 1)
   bool have_aniso = false;
   std::vector<int> taniso;
   std::vector<double> eaniso, kaniso, sb;
   std::vector<double> ext;      // (3,N,M) or empty
   std::vector<double> btorque;  // (3,N,M) or empty
   // llg
   int SDEalgh = 1, mompar = 0;
   double delta_t = 1e-16, temprescale = 1.0;
   std::vector<double> landeg, lambda, temp;  // (N)
   unsigned long long seed = 20261017ull;
   unsigned int ens_offset = 0;   // global index of local ensemble 0 (ensemble sharding)
   bool llg_uniform = true, llg_thermal = false;
   // moments (host copies, original order) used when (re)building layouts
   std::vector<double> h_emom, h_mmom, h_mmom0;
   // layouts
   Layout sd, mc;
   bool sd_built = false, mc_built = false;
   bool lattice_built = false;  // tables live on the device only (asd_build_lattice_table)
   // host tables of a supercell in the reference's atom order (asd_set_lattice_hint): the SD layout is put in brick
   // order like a device-built lattice, so that the same tile / run machinery applies to tables the Fortran host built
   struct { int on = 0, NA = 0, N1 = 0, N2 = 0, N3 = 0, periodic[3] = {0, 0, 0}; } hint;
   bool lat_ordered = false;    // the SD layout of host tables is in brick order (e->lat describes it)
   int state_layout = 0;  // 0 = none, 1 = sd, 2 = mc
   DevBuf<SpinVec> cur, pred;
   DevBuf<double> b2eff, esite, part, red, ring;   // ring: per-sample sums of asd_sd_run
   DevBuf<double> tfield;               // time-dependent uniform field of the steps tf_first .. tf_first + tf_n - 1 ([step][3])
   std::vector<double> h_tfield;        // host copy: the stage launches take the vector of their step as a kernel parameter
   long long tf_first = 0;
   int tf_n = 0;
   double* h_red = nullptr;             // pinned host landing zone of the per-ensemble sums (4 doubles each)
   size_t h_red_n = 0;
   DevBuf<double> msum_part;            // per-tile sums of emomM left by the last corrector launch of asd_sd_steps
   bool msum_fresh = false;
   // moment planes of the MM run kernels: emomM of cur / pred as [M][3][Npad]; mm_valid: mm_cur mirrors `cur` (every writer of
   // `cur` other than the MM stage launches clears it, sd_steps rebuilds the planes when it is false)
   DevBuf<double> mm_cur, mm_pred;
   bool mm_valid = false;
   int msum_ntile = 0;
   DevBuf<double> io_e, io_eM, io_m;   // staging of asd_set_moments / asd_get_moments (kept between calls)
   DevBuf<unsigned int> acc;
   long launches = 0;
   bool committed = false;
   // lattice description (when built on device)
   LatticeDesc lat{};
   Stencil stencil[3];                  // exchange, DM, BQ
   DevBuf<unsigned char> lat_col;       // [Nown] colour of every owned slot (periodic colouring, mc_tile_kernel)
   int lat_ncol = 0, lat_period[3] = {1, 1, 1};
   int mc_layout = -1;                  // -1: default (env ASD_MC_TILES), 0 colour-major, 1 lattice tiles
   // fixed-moment run: frozen[i] != 0 for atoms outside red_atom_list (empty: every atom evolves)
   std::vector<unsigned char> frozen;
   // triangulation for the skyrmion number (asd_set_triangulation): corners as 0-based ORIGINAL atom indices on the host,
   // as slots of one layout on the device (rebuilt when the state moves to the other layout)
   std::vector<int> simp;
   int nsimp = 0, tri_layout = 0;
   DevBuf<int> d_tri;
};

static void launch_cfg(int Npad, int M, dim3& grid, dim3& block) {
   block = dim3(256, 1, 1);
   grid = dim3((Npad + 255) / 256, M, 1);
}

// ------------------------------------------------------------------------------------------------
// Layout construction from host tables
// ------------------------------------------------------------------------------------------------
static int host_orig(asd_engine* e, Layout& L);
static int slab_commit(asd_engine* e);
static int materialise_host_tables(asd_engine* e);
static int slab_push_state(asd_engine* e);
static int lattice_colours(asd_engine* e);
static int fill_lattice_desc(asd_engine* e, int NA, int N1, int N2, int N3l, int N3g, const int* periodic, bool reduced);
static bool has_lattice(const asd_engine* e) { return e->lattice_built || e->lat_ordered; }
static bool mc_block_candidate(const asd_engine* e);
static int mc_block_prepare(asd_engine* e);
static int mc_sweeps_block(asd_engine* e, McParams& p, long nsweeps, long first_sweep);
static int mc_block_visit_order(asd_engine* e, int* order);
template <class K> static void allow_smem(K kernel, size_t bytes);

// staged tile path: gather lists + 16-bit neighbour table (asd_tiles.cuh).  Leaves t.staged = 0 when a tile would
// need more than TILE_UMAX unique slots (layout without locality) or when switched off (ASD_STAGED=0).
static int build_tiles(asd_engine* e, Layout& L, int ts) {
   Tables& t = L.t;
   t.staged = 0; t.ucap = 0; t.ulist = nullptr; t.ucount = nullptr; t.nl16 = nullptr; t.zq8 = (t.z + 7) / 8;
   t.dm16 = nullptr; t.bq16 = nullptr;
   t.tile_slots = ts;
   const char* env = std::getenv("ASD_STAGED");
   if ((env && atoi(env) == 0) || L.is_mc || t.z <= 0 || t.jtens) return 0;
   const long Npad = L.Npad;
   if (t.Nown <= 0) t.Nown = L.Npad;
   const int ntile = (t.Nown + ts - 1) / ts;
   const int* key = L.d_okey.p ? L.d_okey.p : t.orig;
   // lattice layouts: rotate x in the sort key so that periodic images stay next to the tile (asd_tiles.cuh)
   const bool wrap = L.d_okey.p && has_lattice(e) && e->lat.periodic[0] && e->lat.N1 > e->lat.BX && !(std::getenv("ASD_KEYWRAP") && atoi(std::getenv("ASD_KEYWRAP")) == 0);
   const int kna = wrap ? e->lat.NA : 0, kn1 = wrap ? e->lat.N1 : 0, koff = wrap ? (e->lat.N1 - e->lat.BX) / 2 : 0;
   cudaStream_t st = e->stream;
   int r;
   if ((r = L.d_ucount.alloc(ntile))) return r;
   // DM / BQ neighbours join the gather lists (ASD_STAGE_DM=0: keep gathering them from global memory)
   TileExtra x;
   memset(&x, 0, sizeof x);
   const char* sdm = std::getenv("ASD_STAGE_DM");
   if (!(sdm && atoi(sdm) == 0) && ts == 1024) { x.zdm = t.zdm; x.dml = t.dml; x.zbq = t.zbq; x.bql = t.bql; }   // read by llg_runs_kernel<.., 8, .., XS>
   const size_t smem = TILE_BUILD_SMEM;
   CU(cudaFuncSetAttribute(tile_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   tile_gather_kernel<<<ntile, TILE, smem, st>>>(t.Nown, (int)Npad, t.z, t.nl, t.ham, key, 0, 0, L.d_ucount.p, nullptr, nullptr, t.zq8, kna, kn1, koff, ts, x);
   e->launches++;
   CU(cudaGetLastError());
   std::vector<int> cnt(ntile);
   CU(cudaMemcpyAsync(cnt.data(), L.d_ucount.p, (size_t)ntile * sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   const int mx = *std::max_element(cnt.begin(), cnt.end());
   if (mx > TILE_UMAX) return 0;
   const int ucap = ((mx + 31) / 32) * 32;
   if ((r = L.d_ulist.alloc((size_t)ntile * ucap))) return r;
   if ((r = L.d_nl16.alloc((size_t)t.zq8 * Npad))) return r;
   if (x.zdm > 0) { if ((r = L.d_dm16.alloc((size_t)((x.zdm + 7) / 8) * Npad))) return r; x.dm16 = L.d_dm16.p; }
   if (x.zbq > 0) { if ((r = L.d_bq16.alloc((size_t)((x.zbq + 7) / 8) * Npad))) return r; x.bq16 = L.d_bq16.p; }
   tile_gather_kernel<<<ntile, TILE, smem, st>>>(t.Nown, (int)Npad, t.z, t.nl, t.ham, key, 1, ucap, L.d_ucount.p, L.d_ulist.p, L.d_nl16.p, t.zq8, kna, kn1, koff, ts, x);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(st));
   if ((r = L.d_meta.alloc(Npad))) return r;
   zip_meta_kernel<<<(unsigned)((Npad + 255) / 256), 256, 0, st>>>((int)Npad, t.ham, t.orig, L.d_meta.p);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(st));
   t.meta = L.d_meta.p;
   t.staged = 1; t.ucap = ucap; t.ulist = L.d_ulist.p; t.ucount = L.d_ucount.p; t.nl16 = L.d_nl16.p;
   t.dm16 = x.dm16; t.bq16 = x.bq16;
   return 0;
}

// run-compressed table (asd_runs.cuh): groups of 4 x-runs -> union rows.  Leaves t.runs = 0 unless every group of the
// layout is regular.
static int build_runs(asd_engine* e, Layout& L) {
   constexpr int R = 4;
   Tables& t = L.t;
   t.runs = 0; t.urow = 0; t.utab = nullptr; t.mm = 0;
   if (!t.staged || !L.reduced || !t.cpl_param || L.is_mc || !has_lattice(e)) return 0;
   if (t.z * R >= RUN_MAXPAIR) return 0;
   const int ngroup = (t.Nown + R * 32 - 1) / (R * 32);
   const int ntile = (t.Nown + t.tile_slots - 1) / t.tile_slots;
   const int galloc = ntile * (t.tile_slots / (R * 32));      // whole tiles: the stage kernel copies NW rows per tile
   cudaStream_t st = e->stream;
   int r;
   if ((r = L.d_gcount.alloc(ngroup))) return r;
   // MM instantiations of the run kernel (gather list staged from the moment planes with cp.async): 1024-slot tiles whose list fits
   // the fixed plane stride, none of the XS cases (DM / BQ positions, short lists); a slab pushes the planes of its boundary atoms
   // into the neighbours' halos next to the spins
   {
      const char* menv = std::getenv("ASD_MM");
      const bool xs = t.dm16 != nullptr || t.bq16 != nullptr || t.ucap <= 6 * 256;
      // ... and only plain Heisenberg layouts: paired with the LEAN integrator loop the planes give 0.456 -> 0.429 ms per step at
      // bcc 128^3, with the general loop 0.468 (measured, profiles/README: the general instantiation is at its register limit)
      const bool plain = t.zdm == 0 && t.zbq == 0 && !t.jtens;
      t.mm = (!(menv && atoi(menv) == 0) && t.tile_slots == 1024 && t.ucap + 32 <= MM_PLANE - 32 && !xs && plain) ? 1 : 0;
   }
   const unsigned pos_scale = t.mm ? 8u : 24u;
   const int pad = t.mm ? WALK_U : 1;                       // union rows of MM layouts: mask classes padded to the walk's unroll
   const unsigned null_pos = (unsigned)(MM_PLANE - 32);     // the zero moment the MM kernel keeps for the null entries
   run_union_kernel<R><<<ngroup, 32, 0, st>>>(t.Nown, L.Npad, t.z, t.nl16, t.meta, t.lsize, 0, 0, L.d_gcount.p, nullptr, pos_scale, pad, null_pos);
   e->launches++;
   CU(cudaGetLastError());
   std::vector<int> cnt(ngroup);
   CU(cudaMemcpyAsync(cnt.data(), L.d_gcount.p, (size_t)ngroup * sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   int mx = 0, mxu = 0;
   for (int c : cnt) {
      if (c < 0) {
         if (std::getenv("ASD_DEBUG")) {
            int hist[6] = {0, 0, 0, 0, 0, 0};
            for (int q : cnt) if (q < 0 && q >= -5) hist[-q]++;
            fprintf(stderr, "[asd] run table refused: %d groups; not a lane prefix %d, mixed rows %d, too many pairs %d, split neighbour run %d, duplicate neighbour %d\n",
                    ngroup, hist[1], hist[2], hist[3], hist[4], hist[5]);
         }
         t.mm = 0;
         return 0;
      }
      mx = std::max(mx, c >> 10);      // entries with the padding of the mask classes
      mxu = std::max(mxu, c & 1023);   // distinct neighbour runs
   }
   if (mx == 0 || mx > 255) { t.mm = 0; return 0; }
   const int urow = 1 + mx + pad;   // header word + entries + `pad` spare (zero) entries
   if ((r = L.d_utab.alloc((size_t)galloc * urow))) return r;
   CU(cudaMemsetAsync(L.d_utab.p, 0, (size_t)galloc * urow * sizeof(uint4), st));
   run_union_kernel<R><<<ngroup, 32, 0, st>>>(t.Nown, L.Npad, t.z, t.nl16, t.meta, t.lsize, 1, urow, L.d_gcount.p, L.d_utab.p, pos_scale, pad, null_pos);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(st));
   t.union_max = mxu;
   t.runs = R; t.urow = urow; t.utab = L.d_utab.p;
   return 0;
}

// external_field(3,N,M) of the engine -> layout: a uniform field rides in the kernel parameters, anything else
// is permuted into device order.  Callable on a committed layout (the drivers switch between the initial-phase and
// the measurement-phase field, sd_driver.f90:122 / :406).
static int apply_external_field(asd_engine* e, Layout& L) {
   const int N = e->N, M = e->M;
   Tables& t = L.t;
   t.ext_uniform = 1; t.hext[0] = t.hext[1] = t.hext[2] = 0.0; t.ext = nullptr;
   if (e->ext.empty()) return 0;
   bool uni = true;
   for (size_t q = 0; q < (size_t)N * M && uni; q++)
      for (int a = 0; a < 3; a++) if (e->ext[3 * q + a] != e->ext[a]) { uni = false; break; }
   if (uni) { for (int a = 0; a < 3; a++) t.hext[a] = e->ext[a]; return 0; }
   int r = host_orig(e, L);
   if (r) return r;
   const long Npad = L.Npad;
   std::vector<double> h((size_t)M * 3 * Npad, 0.0);
   for (int k = 0; k < M; k++)
      for (long s = 0; s < Npad; s++) {
         const int o = L.orig[s];
         if (o < 0) continue;
         for (int a = 0; a < 3; a++) h[((size_t)k * 3 + a) * Npad + s] = e->ext[a + 3 * ((size_t)o + (size_t)N * k)];
      }
   if ((r = L.d_ext.upload(h, e->stream))) return r;
   t.ext_uniform = 0; t.ext = L.d_ext.p;
   return 0;
}

// shared tail of layout construction: shared-memory plan, per-atom arrays (anisotropy, fields) in device order
static int finish_layout(asd_engine* e, Layout& L) {
   const int N = e->N, NH = e->NH, M = e->M;
   const long Npad = L.Npad;
   cudaStream_t st = e->stream;
   Tables& t = L.t;
   int r;
   // ---- shared-memory staging plan for reduced couplings ----
   L.smem_bytes = 0; t.sm_cp = t.sm_dm = t.sm_bq = 0;
   if (L.reduced) {
      size_t n0 = (size_t)NH * t.z * (t.jtens ? 9 : 1), n1 = (size_t)NH * t.zdm * 3, n2 = (size_t)NH * t.zbq;
      if ((n0 + n1 + n2) * 8 <= 40 * 1024) { t.sm_cp = (int)n0; t.sm_dm = (int)n1; t.sm_bq = (int)n2; L.smem_bytes = (n0 + n1 + n2) * 8; }
   }
   // ---- field path of the stage kernels: run kernel on big tiles > staged tiles > direct gathers
   //      (experiment knobs: ASD_VARIANT, ASD_PF, ASD_STAGED, ASD_RUNS = 0 | 256 | 512 | 1024) ----
   {
      const char* var = std::getenv("ASD_VARIANT");
      // tensorial exchange: direct gathers in site_field, none of the scalar-coupling fast paths
      const int variant = t.jtens ? 0 : (var ? atoi(var) : 3);
      t.nl4 = nullptr; t.cp4 = nullptr; t.zq = (t.z + 3) / 4; t.pf_tiles = 0; t.cpl_param = 0;
      t.runs = 0; t.urow = 0; t.utab = nullptr;
      if (variant >= 3 && L.reduced && t.z > 0 && (size_t)NH * t.z <= 256) {
         std::vector<double> rows((size_t)NH * t.z);
         CU(cudaMemcpy(rows.data(), t.cp, rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
         for (size_t q = 0; q < rows.size(); q++) t.cpl_small[q] = rows[q];
         t.cpl_param = 1;
      }
      // tile size of the run kernel: the whole super-brick by default
      int big = (has_lattice(e) && !L.is_mc) ? e->lat.NA * e->lat.P * e->lat.SY * e->lat.SZ : 256;   // slots of a super-brick
      // 1024-slot tiles when they divide the super-brick (bcc: the super-brick itself; fcc, four basis atoms: half of it); smaller
      // tiles only on request (ASD_RUNS): measured slower than the staged kernel.  (Short lists, fcc z = 18: the run kernel on
      // the moment planes 0.347 ms per step at 128 x 128 x 64 x 4, the staged one-atom-per-thread kernel 0.377.)
      big = (big >= 1024 && big % 1024 == 0) ? 1024 : 256;
      const char* renv = std::getenv("ASD_RUNS");
      if (renv) big = atoi(renv);
      if (big != 0 && big != 256 && big != 512 && big != 1024) return fail(-1, "ASD_RUNS must be 0, 256, 512 or 1024");
      if (big > 256 && (!has_lattice(e) || L.is_mc || (e->lat.NA * e->lat.P * e->lat.SY * e->lat.SZ) % big != 0)) big = 256;
      if (big >= 256 && (renv || big > 256)) {
         if ((r = build_tiles(e, L, big))) return r;
         if ((r = build_runs(e, L))) return r;
      }
      if (!t.runs && (r = build_tiles(e, L, 256))) return r;
      // staged layouts use nl16 / the union rows in the LLG kernels; the field-only / MC kernels read the plain nl table
      if (variant >= 3 && t.z > 0) {
         L.d_nl4.release(); L.d_cp4.release();
         if (!t.staged && (r = L.d_nl4.alloc((size_t)t.zq * Npad))) return r;
         if (!L.reduced && (r = L.d_cp4.alloc((size_t)t.zq * Npad))) return r;
         if (L.d_nl4.p || L.d_cp4.p) {
            vectorise_table_kernel<<<dim3((unsigned)((Npad + 255) / 256), t.zq), 256, 0, st>>>((int)Npad, t.z, t.zq, t.nl, L.reduced ? nullptr : t.cp,
                                                                                    L.d_nl4.p, L.reduced ? nullptr : L.d_cp4.p);
            e->launches++;
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(st));
         }
         t.nl4 = L.d_nl4.p; t.cp4 = L.d_cp4.p;
         const char* pf = std::getenv("ASD_PF");
         int sms = 148;
         cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
         t.pf_tiles = pf ? atoi(pf) : (t.runs ? sms : sms * 3);   // prefetch distance in tiles (measured optimum)
      }
   }
   // ---- per-atom arrays in device order ----
   auto permute = [&](const std::vector<double>& src, int ncomp, bool per_ens, DevBuf<double>& dst) -> int {
      int rr = host_orig(e, L);
      if (rr) return rr;
      const int K = per_ens ? M : 1;
      std::vector<double> h((size_t)K * ncomp * Npad, 0.0);
      for (int k = 0; k < K; k++)
         for (long s = 0; s < Npad; s++) {
            const int o = L.orig[s];
            if (o < 0) continue;
            for (int a = 0; a < ncomp; a++) h[((size_t)k * ncomp + a) * Npad + s] = src[a + (size_t)ncomp * ((size_t)o + (per_ens ? (size_t)N * k : 0))];
         }
      return dst.upload(h, st);
   };
   if (e->have_aniso) {
      if ((r = host_orig(e, L))) return r;
      std::vector<int> ta(Npad, 0);
      for (long s = 0; s < Npad; s++) if (L.orig[s] >= 0) ta[s] = e->taniso[L.orig[s]];
      if ((r = L.d_taniso.upload(ta, st))) return r;
      if ((r = permute(e->eaniso, 3, false, L.d_eaniso))) return r;
      if ((r = permute(e->kaniso, 2, false, L.d_kaniso))) return r;
      if ((r = permute(e->sb, 1, false, L.d_sb))) return r;
      t.do_aniso = 1; t.taniso = L.d_taniso.p; t.eaniso = L.d_eaniso.p; t.kaniso = L.d_kaniso.p; t.sb = L.d_sb.p;
      // same anisotropy on every atom of a Hamiltonian row (the usual case of do_reduced Y): constant bank, no loads
      t.aniso_rows = 0;
      if (L.reduced && NH <= 8) {
         std::vector<int> first(NH, -1);
         bool uni = true;
         for (int i = 0; i < N && uni; i++) {
            const int h = e->aHam[i] - 1;
            if (first[h] < 0) { first[h] = i; continue; }
            const int f = first[h];
            uni = e->taniso[i] == e->taniso[f] && e->kaniso[2 * (size_t)i] == e->kaniso[2 * (size_t)f] &&
                  e->kaniso[2 * (size_t)i + 1] == e->kaniso[2 * (size_t)f + 1] && e->sb[i] == e->sb[f] &&
                  e->eaniso[3 * (size_t)i] == e->eaniso[3 * (size_t)f] && e->eaniso[3 * (size_t)i + 1] == e->eaniso[3 * (size_t)f + 1] &&
                  e->eaniso[3 * (size_t)i + 2] == e->eaniso[3 * (size_t)f + 2];
         }
         if (uni) {
            for (int h = 0; h < NH; h++) {
               const int f = std::max(first[h], 0);
               double* a = t.aniso_small[h];
               a[0] = (double)e->taniso[f]; a[1] = e->kaniso[2 * (size_t)f]; a[2] = e->kaniso[2 * (size_t)f + 1];
               a[3] = e->eaniso[3 * (size_t)f]; a[4] = e->eaniso[3 * (size_t)f + 1]; a[5] = e->eaniso[3 * (size_t)f + 2];
               a[6] = e->sb[f]; a[7] = 0.0;
            }
            t.aniso_rows = 1;
         }
      }
   }
   if ((r = apply_external_field(e, L))) return r;
   if (!e->btorque.empty()) { if ((r = permute(e->btorque, 3, true, L.d_btorque))) return r; t.btorque = L.d_btorque.p; }
   return 0;
}

// greedy colouring of the symmetrised union of the neighbour tables (host, O(N z log z))
static int colour_graph(const asd_engine* e, std::vector<int>& colour) {
   const int N = e->N;
   const HostTable* T[3] = {&e->ex, &e->dm, &e->bq};
   int ztot = 0;
   for (auto* t : T) if (t->present()) ztot += t->z;
   // adj[i*ztot .. ) = sorted neighbours of i (0-based, self and empty entries removed), deg[i] of them
   std::vector<int> adj((size_t)N * ztot), deg(N, 0);
   for (int i = 0; i < N; i++) {
      int* a = adj.data() + (size_t)i * ztot;
      int n = 0;
      for (auto* t : T) {
         if (!t->present()) continue;
         for (int j = 0; j < t->z; j++) {
            const int nb = t->list[(size_t)j + (size_t)t->z * i];
            if (nb > 0 && nb - 1 != i) a[n++] = nb - 1;
         }
      }
      std::sort(a, a + n);
      n = (int)(std::unique(a, a + n) - a);
      deg[i] = n;
   }
   // lists are symmetric for every physical table; add the reverse of any edge that is not (e.g. DM maps, sym 0)
   std::vector<std::vector<int>> rev(N);
   for (int i = 0; i < N; i++) {
      const int* a = adj.data() + (size_t)i * ztot;
      for (int q = 0; q < deg[i]; q++) {
         const int nb = a[q];
         const int* b = adj.data() + (size_t)nb * ztot;
         if (!std::binary_search(b, b + deg[nb], i)) rev[nb].push_back(i);
      }
   }
   colour.assign(N, -1);
   std::vector<int> mark;
   int ncol = 0;
   for (int i = 0; i < N; i++) {
      mark.assign(ncol + 1, 0);
      auto see = [&](int nb) { int c = colour[nb]; if (c >= 0 && c <= ncol) mark[c] = 1; };
      const int* a = adj.data() + (size_t)i * ztot;
      for (int q = 0; q < deg[i]; q++) see(a[q]);
      for (int nb : rev[i]) see(nb);
      int c = 0;
      while (c < ncol && mark[c]) c++;
      colour[i] = c;
      if (c == ncol) ncol++;
   }
   return ncol;
}

static int build_layout(asd_engine* e, Layout& L, bool colour_major) {
   const int N = e->N, NH = e->NH, M = e->M;
   e->tri_layout = 0;          // slots change: the device copy of the triangulation is rebuilt on its next use
   L.d_frozen.release();       // ... and so are the per-site arrays kept in device order (fill_llg re-permutes them)
   L.d_lambda.release(); L.d_landeg.release(); L.d_temp.release();
   L.N = N; L.NH = NH; L.M = M;
   L.reduced = NH < N;
   L.is_mc = colour_major;
   // ---- ordering: groups = (colour, ham row) for MC, (ham row) for SD; stable within a group.  SD layout of a
   //      supercell whose shape the host announced (asd_set_lattice_hint): brick order, like a device-built lattice ----
   std::vector<int> colour;
   int ncol = 1;
   if (colour_major) ncol = colour_graph(e, colour);
   const int nrow = L.reduced ? NH : 1;
   bool brick = false;
   if (!colour_major) {
      e->lat_ordered = false;
      const auto& h = e->hint;
      brick = h.on && !e->lattice_built && (long)h.NA * h.N1 * h.N2 * h.N3 == N && (!L.reduced || NH == h.NA);
      if (brick && L.reduced)
         for (int i = 0; i < N && brick; i++) if (e->aHam[i] != i % h.NA + 1) brick = false;
      const char* env = std::getenv("ASD_HINT");
      if (env && atoi(env) == 0) brick = false;
   }
   std::vector<long> gstart;
   if (brick) {
      const auto& h = e->hint;
      int r = fill_lattice_desc(e, h.NA, h.N1, h.N2, h.N3, h.N3, h.periodic, L.reduced);
      if (r) return r;
      const LatticeDesc& d = e->lat;
      L.Npad = d.Npad;
      L.orig.assign(d.Npad, -1);
      L.slot_of.assign(N, -1);
      for (int i = 0; i < N; i++) {
         const int i0 = i % h.NA, c = i / h.NA;
         const int s = lattice_slot(d, i0, c % h.N1, (c / h.N1) % h.N2, c / (h.N1 * h.N2));
         L.orig[s] = i; L.slot_of[i] = s;
      }
      e->lat_ordered = true;
   } else {
      std::vector<long> gcount((size_t)ncol * nrow, 0);
      auto group_of = [&](int i) { return (size_t)(colour_major ? colour[i] : 0) * nrow + (L.reduced ? e->aHam[i] - 1 : 0); };
      for (int i = 0; i < N; i++) gcount[group_of(i)]++;
      gstart.assign(gcount.size() + 1, 0);
      for (size_t g = 0; g < gcount.size(); g++) gstart[g + 1] = gstart[g] + ((gcount[g] + 31) / 32) * 32;
      const long Npad = gstart.back();
      if (Npad > 2000000000L) return fail(-3, "too many atoms for 32-bit device indices");
      L.Npad = (int)Npad;
      L.orig.assign(Npad, -1);
      L.slot_of.assign(N, -1);
      std::vector<long> fill(gstart.begin(), gstart.end() - 1);
      for (int i = 0; i < N; i++) { long s = fill[group_of(i)]++; L.orig[s] = i; L.slot_of[i] = (int)s; }
   }
   const long Npad = L.Npad;
   L.colour_first.clear(); L.colour_count.clear();
   if (colour_major)
      for (int c = 0; c < ncol; c++) {
         L.colour_first.push_back((int)gstart[(size_t)c * nrow]);
         L.colour_count.push_back((int)(gstart[(size_t)(c + 1) * nrow] - gstart[(size_t)c * nrow]));
      }
   std::vector<int> ham(Npad, -1);
   for (long s = 0; s < Npad; s++) if (L.orig[s] >= 0) ham[s] = L.reduced ? e->aHam[L.orig[s]] - 1 : 0;
   cudaStream_t st = e->stream;
   int r;
   if ((r = L.d_ham.upload(ham, st))) return r;
   if ((r = L.d_orig.upload(L.orig, st))) return r;
   if (brick) { if ((r = L.d_okey.upload(L.orig, st))) return r; }   // sort key of the gather lists = atom index (x fastest)
   else L.d_okey.release();
   Tables& t = L.t;
   memset(&t, 0, sizeof t);
   t.N = N; t.Npad = L.Npad; t.Nown = L.Npad; t.M = M; t.NH = NH; t.reduced = L.reduced ? 1 : 0;
   t.ens_offset = e->ens_offset;
   t.ham = L.d_ham.p; t.orig = L.d_orig.p;
   // ---- pair tables ----
   auto do_table = [&](const HostTable& T, DevBuf<int>& d_list, DevBuf<double>& d_coup, DevBuf<int>& d_size, const char* name) -> int {
      if (!T.present()) return 0;
      const int z = T.z, nc = T.ncomp;
      std::vector<int> nl((size_t)z * Npad);
      for (long s = 0; s < Npad; s++) {
         const int o = L.orig[s];
         for (int j = 0; j < z; j++) nl[(size_t)j * Npad + s] = (int)s;  // padding -> self
         if (o < 0) continue;
         const int row = L.reduced ? e->aHam[o] - 1 : o;
         const int n = T.lsize[row];
         for (int j = 0; j < n; j++) {
            const int nb = T.list[(size_t)j + (size_t)z * o];
            if (nb < 1 || nb > N)
               return fail(-4, "%s table: atom %d has no neighbour in slot %d although nlistsize(aHam)=%d "
                               "(do_reduced needs complete neighbour sets on every atom)", name, o + 1, j + 1, n);
            nl[(size_t)j * Npad + s] = L.slot_of[nb - 1];
         }
      }
      int rr;
      if ((rr = d_list.upload(nl, st))) return rr;
      if (L.reduced) {
         // [NH][z][ncomp]: row-major copy of coup(ncomp,z,NH)
         std::vector<double> c((size_t)NH * z * nc);
         for (int h = 0; h < NH; h++)
            for (int j = 0; j < z; j++)
               for (int a = 0; a < nc; a++) c[((size_t)h * z + j) * nc + a] = T.coup[a + (size_t)nc * (j + (size_t)z * h)];
         if ((rr = d_coup.upload(c, st))) return rr;
         if ((rr = d_size.upload(T.lsize, st))) return rr;
      } else {
         // [ncomp][z][Npad], zero beyond nlistsize
         std::vector<double> c((size_t)nc * z * Npad, 0.0);
         for (long s = 0; s < Npad; s++) {
            const int o = L.orig[s];
            if (o < 0) continue;
            const int n = T.lsize[o];
            for (int j = 0; j < n; j++)
               for (int a = 0; a < nc; a++) c[((size_t)a * z + j) * Npad + s] = T.coup[a + (size_t)nc * (j + (size_t)z * o)];
         }
         if ((rr = d_coup.upload(c, st))) return rr;
      }
      return 0;
   };
   if ((r = do_table(e->ex, L.d_nl, L.d_cp, L.d_lsize, "exchange"))) return r;
   if ((r = do_table(e->dm, L.d_dml, L.d_dmv, L.d_dmsize, "DM"))) return r;
   if ((r = do_table(e->bq, L.d_bql, L.d_jbq, L.d_bqsize, "BQ"))) return r;
   t.z = e->ex.z; t.nl = L.d_nl.p; t.cp = L.d_cp.p; t.lsize = L.d_lsize.p;
   t.jtens = e->jtensor ? 1 : 0;
   t.zdm = e->dm.z; t.dml = L.d_dml.p; t.dmv = L.d_dmv.p; t.dmsize = L.d_dmsize.p;
   t.zbq = e->bq.z; t.bql = L.d_bql.p; t.jbq = L.d_jbq.p; t.bqsize = L.d_bqsize.p;
   L.zs[0] = e->ex.z; L.zs[1] = e->dm.z; L.zs[2] = e->bq.z;
   L.d_nlrow.release();
   L.d_classes.release();
   L.d_cprow.release();
   // atom-major copies for the cooperative colour kernels (small colour classes / long lists); per-atom couplings (do_reduced N,
   // random alloys) up to 2e8 table entries (1.6 GB of couplings)
   if (colour_major && !e->jtensor && (L.reduced || (size_t)Npad * e->ex.z <= (size_t)200000000)) {
      const int z = e->ex.z;
      std::vector<int> rowm((size_t)Npad * z);
      std::vector<double> rowc;
      if (!L.reduced) rowc.assign((size_t)Npad * z, 0.0);
      for (long s = 0; s < Npad; s++) {
         const int o = L.orig[s];
         for (int j = 0; j < z; j++) {
            int v = (int)s;
            if (o >= 0 && j < e->ex.lsize[e->aHam[o] - 1]) {
               v = L.slot_of[e->ex.list[(size_t)j + (size_t)z * o] - 1];
               if (!L.reduced) rowc[(size_t)s * z + j] = e->ex.coup[(size_t)j + (size_t)z * (e->aHam[o] - 1)];
            }
            rowm[(size_t)s * z + j] = v;
         }
      }
      if ((r = L.d_nlrow.upload(rowm, st))) return r;
      if (!L.reduced && (r = L.d_cprow.upload(rowc, st))) return r;
   }
   r = finish_layout(e, L);
   L.t.nlrow = L.d_nlrow.p;
   L.t.cprow = L.d_cprow.p;
   return r;
}

// per-site LLG parameter arrays of a layout (uniform -> scalars)
static int fill_llg(asd_engine* e, Layout& L, LlgParams& p, unsigned long long step) {
   memset(&p, 0, sizeof p);
   const bool uni = e->llg_uniform;
   p.per_site = uni ? 0 : 1;
   p.landeg = e->landeg.empty() ? 1.0 : e->landeg[0];
   p.lambda = e->lambda.empty() ? 0.05 : e->lambda[0];
   p.temp = e->temp.empty() ? 0.0 : e->temp[0];
   if (!uni && L.d_lambda.p == nullptr) {
      auto perm = [&](const std::vector<double>& src, DevBuf<double>& dst) -> int {
         int rr = host_orig(e, L);
         if (rr) return rr;
         std::vector<double> h(L.Npad, 0.0);
         for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) h[s] = src[L.orig[s]];
         return dst.upload(h, e->stream);
      };
      int r;
      if ((r = perm(e->landeg, L.d_landeg))) return r;
      if ((r = perm(e->lambda, L.d_lambda))) return r;
      if ((r = perm(e->temp, L.d_temp))) return r;
   }
   p.landeg_a = L.d_landeg.p; p.lambda_a = L.d_lambda.p; p.temp_a = L.d_temp.p;
   if (!e->frozen.empty() && L.d_frozen.p == nullptr) {
      int rr = host_orig(e, L);
      if (rr) return rr;
      std::vector<unsigned char> h(L.Npad, 0);
      for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) h[s] = e->frozen[L.orig[s]];
      if ((rr = L.d_frozen.upload(h, e->stream))) return rr;
   }
   p.frozen = e->frozen.empty() ? nullptr : L.d_frozen.p;
   p.tfield = e->tf_n > 0 ? e->tfield.p : nullptr; p.tf_first = e->tf_first; p.tf_n = e->tf_n;
   p.delta_t = e->delta_t; p.gamma = e->gamma; p.k_bolt = e->k_bolt; p.mub = e->mub; p.temprescale = e->temprescale;
   p.mompar = e->mompar; p.mmom0 = L.d_mmom0.p;
   p.seed = e->seed; p.step = step;
   p.thermal = e->llg_thermal ? 1 : 0;
   p.mm_cur = (L.t.runs && L.t.mm) ? e->mm_cur.p : nullptr;
   p.mm_pred = (L.t.runs && L.t.mm) ? e->mm_pred.p : nullptr;
   {
      // same expressions, same order as the per-site branch of llg_stage_kernel (volatile: no host-side contraction)
      volatile double lam = p.lambda, one = 1.0;
      volatile double lam2 = lam * lam;
      volatile double den = one + lam2;
      p.u_lldamp = one / den;
      volatile double dt0 = p.delta_t * one;
      volatile double dt1 = dt0 * p.gamma;
      p.u_dt = dt1 * p.u_lldamp;
      p.u_sqrtdt = std::sqrt(p.u_dt);
      volatile double d0 = lam / den;
      volatile double d1 = d0 * p.k_bolt;
      volatile double d2 = d1 / p.gamma;
      volatile double d3 = d2 / p.mub;
      volatile double gg = p.gamma / one;
      p.u_Dk = d3 * gg;
      volatile double n0 = 2.0 * lam;
      volatile double n1 = n0 * p.k_bolt;
      volatile double q0 = p.delta_t * p.gamma;
      volatile double q1 = q0 * p.mub;
      p.u_Dp = n1 / q1;
      volatile double s0 = 2.0 * p.u_Dk;
      volatile double s1 = s0 * p.temp;
      volatile double s2 = s1 * p.temprescale;
      p.u_sig1 = std::sqrt(s2);
      volatile double r0 = p.u_Dp * p.temprescale;
      volatile double r1 = r0 * p.temp;
      p.u_sig5 = std::sqrt(r1);
   }
   return 0;
}

// ------------------------------------------------------------------------------------------------
// state movement
// ------------------------------------------------------------------------------------------------
// host arrays (Fortran shapes, original atom order) -> packed device order.  emom / mmom may point at the caller's
// buffers (pinned or pageable): they are copied with cudaMemcpyAsync on the engine's stream and packed on the device.
static int upload_state_from(asd_engine* e, Layout& L, const double* emom, const double* mmom, const double* mmom0) {
   const size_t NM = (size_t)e->N * e->M;
   e->msum_fresh = false;
   e->mm_valid = false;
   int r;
   if ((r = e->io_e.alloc(3 * NM))) return r;
   if ((r = e->io_m.alloc(NM))) return r;
   CU(cudaMemcpyAsync(e->io_e.p, emom, 3 * NM * sizeof(double), cudaMemcpyHostToDevice, e->stream));
   CU(cudaMemcpyAsync(e->io_m.p, mmom, NM * sizeof(double), cudaMemcpyHostToDevice, e->stream));
   if ((r = e->cur.alloc((size_t)L.Npad * e->M))) return r;
   if ((r = e->pred.alloc((size_t)L.Npad * e->M))) return r;
   dim3 g, b;
   launch_cfg(L.Npad, e->M, g, b);
   // only the owned slots: the halo slots of a slab belong to the neighbours' pushes
   pack_kernel<<<g, b, 0, e->stream>>>(e->N, L.t.Nown, L.Npad, e->M, L.d_orig.p, e->io_e.p, e->io_m.p, e->cur.p);
   pack_kernel<<<g, b, 0, e->stream>>>(e->N, L.t.Nown, L.Npad, e->M, L.d_orig.p, e->io_e.p, e->io_m.p, e->pred.p);
   e->launches += 2;
   CU(cudaGetLastError());
   if (e->mompar != 0) {
      if ((r = host_orig(e, L))) return r;
      std::vector<double> h((size_t)e->M * L.Npad, 0.0);
      const double* src = mmom0 ? mmom0 : mmom;
      for (int k = 0; k < e->M; k++)
         for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) h[(size_t)k * L.Npad + s] = src[(size_t)L.orig[s] + (size_t)e->N * k];
      if ((r = L.d_mmom0.upload(h, e->stream))) return r;
   }
   CU(cudaStreamSynchronize(e->stream));
   return slab_push_state(e);
}

static int upload_state(asd_engine* e, Layout& L) {
   return upload_state_from(e, L, e->h_emom.data(), e->h_mmom.data(), e->h_mmom0.empty() ? nullptr : e->h_mmom0.data());
}

// A halo wait that timed out (halo_wait_kernel) leaves the device flag sb.err set: every entry point that synchronises the
// stream reads it, so that a lost or slow peer surfaces as an error code instead of silently stale halos.
static int slab_check(asd_engine* e) {
   Slab& sb = e->slab;
   if (!sb.on || !sb.err.p) return 0;
   int flag = 0;
   CU(cudaMemcpyAsync(&flag, sb.err.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   if (flag) return fail(-12, "slab: timed out waiting for the halo of the %s neighbour; the state is not valid", flag == 1 ? "lower" : "upper");
   return 0;
}

static int download_state(asd_engine* e, Layout& L, double* emom, double* emomM, double* mmom) {
   const size_t NM = (size_t)e->N * e->M;
   int r;
   if (emom && (r = e->io_e.alloc(3 * NM))) return r;
   if (emomM && (r = e->io_eM.alloc(3 * NM))) return r;
   if (mmom && (r = e->io_m.alloc(NM))) return r;
   dim3 g, b;
   launch_cfg(L.Npad, e->M, g, b);
   unpack_kernel<<<g, b, 0, e->stream>>>(e->N, L.Npad, e->M, L.d_orig.p, e->cur.p, emom ? e->io_e.p : nullptr,
                                         emomM ? e->io_eM.p : nullptr, mmom ? e->io_m.p : nullptr);
   e->launches++;
   CU(cudaGetLastError());
   if (emom) CU(cudaMemcpyAsync(emom, e->io_e.p, 3 * NM * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   if (emomM) CU(cudaMemcpyAsync(emomM, e->io_eM.p, 3 * NM * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   if (mmom) CU(cudaMemcpyAsync(mmom, e->io_m.p, NM * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   return slab_check(e);
}

// the device holds the only copy of the state (direct upload): bring it back before a layout is rebuilt
static int stash_state_to_host(asd_engine* e) {
   if (e->state_layout == 0 || !e->h_emom.empty()) return 0;
   Layout& from = (e->state_layout == 1) ? e->sd : e->mc;
   e->h_emom.resize(3 * (size_t)e->N * e->M);
   e->h_mmom.resize((size_t)e->N * e->M);
   return download_state(e, from, e->h_emom.data(), nullptr, e->h_mmom.data());
}

// make sure the spin buffers are in layout `want` (1 sd, 2 mc); converts through the host copy when switching
static int ensure_layout(asd_engine* e, int want) {
   if (!e->committed) return fail(-2, "asd_commit has not been called");
   if (want == 2 && !e->mc_built) {
      if (e->slab.on) return fail(-5, "internal: a slab runs Monte Carlo on the lattice layout");
      if (e->lattice_built) {
         // the colour-major layout is built on the host: bring the device-built tables back in the reference's shape
         if ((long)e->N > 40000000L) return fail(-5, "Monte Carlo layout of a device-built lattice is limited to 4e7 atoms per engine");
         int r = materialise_host_tables(e);
         if (r) return r;
      }
      int r = build_layout(e, e->mc, true);
      if (r) return r;
      e->mc_built = true;
   }
   if (e->state_layout == want) return 0;
   if (e->state_layout != 0) {
      Layout& from = (e->state_layout == 1) ? e->sd : e->mc;
      e->h_emom.resize(3 * (size_t)e->N * e->M);
      e->h_mmom.resize((size_t)e->N * e->M);
      int r = download_state(e, from, e->h_emom.data(), nullptr, e->h_mmom.data());
      if (r) return r;
   }
   if (e->h_emom.empty()) return fail(-6, "moments have not been set (asd_set_moments)");
   Layout& to = (want == 1) ? e->sd : e->mc;
   int r = upload_state(e, to);
   if (r) return r;
   e->state_layout = want;
   return 0;
}

// ------------------------------------------------------------------------------------------------
// compute
// ------------------------------------------------------------------------------------------------
template <class K>
static void allow_smem(K kernel, size_t bytes) {
   // opt in to > 48 KB of dynamic shared memory: the attribute is per device and per kernel (kernels of one signature
   // share K, engines of one process may sit on different devices)
   static std::map<std::pair<int, const void*>, size_t> granted;
   if (bytes <= 48 * 1024) return;
   int dev = 0;
   cudaGetDevice(&dev);
   size_t& g = granted[std::make_pair(dev, (const void*)kernel)];
   if (bytes > g) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      g = bytes;
   }
}

// launch with or without the programmatic-dependent-launch attribute (the kernel must issue griddepcontrol.wait before it
// touches anything the previous kernel of the stream writes)
template <class K, class... A>
static void launch_dep(bool pdl, K kernel, dim3 g, dim3 b, size_t smem, cudaStream_t st, A... args) {
   if (!pdl) { kernel<<<g, b, smem, st>>>(args...); return; }
   cudaLaunchConfig_t cfg;
   memset(&cfg, 0, sizeof cfg);
   cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = st;
   cudaLaunchAttribute at[1];
   at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   at[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = at; cfg.numAttrs = 1;
   cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int SOLVER, int STAGE, bool EDGE, bool MSUM>
static void launch_stage_range2(asd_engine* e, Layout& L, const LlgParams& p, const EdgeParams& ep, const TileRange& tr, int ntiles) {
   dim3 g(ntiles, e->M, 1);
   const dim3 b(256, 1, 1);
   // fixed-moment runs (asd_set_evolving_atoms): only the direct one-atom-per-thread kernel is compiled with the frozen mask;
   // it works on 256-slot tiles whatever tile size the layout was built for (never on a slab: refused at the setter)
   const bool fr = p.frozen != nullptr;
   TileRange trd = tr;
   if (fr && L.t.tile_slots != 256) {
      const int n256 = (L.t.Nown + 255) / 256;
      g.x = (unsigned)n256;
      trd = TileRange{0, n256, 0};
   }
   // Dependent launches pay when a grid runs for more than one wave (the next stage's CTAs fill the SMs the last wave leaves
   // idle): measured +1.3 % at 128^3, +6.5 % at 64^3, but -14 % for a grid that fits the GPU at once (32^3: 256 CTAs), so a
   // grid smaller than the resident capacity is launched the ordinary way.
   static const bool pdl_all = !(std::getenv("ASD_PDL") && atoi(std::getenv("ASD_PDL")) == 0);
   int sms_ = 148;
   cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, e->device);
   const bool big_grid = (long)g.x * g.y > (long)sms_ * ((L.t.runs && !fr) ? 2 : 4);
   const bool pdl_s = pdl_all && !EDGE && !e->slab.on && big_grid;
   if (L.t.runs && !fr) {
      const int NW = L.t.tile_slots / 128;
      const bool mm = L.t.mm && p.mm_cur != nullptr;
      const size_t smem = (size_t)((L.t.sm_dm + L.t.sm_bq + 1) & ~1) * sizeof(double) + (size_t)3 * (mm ? MM_PLANE : (L.t.ucap + 32)) * sizeof(double) +
                          (size_t)NW * L.t.urow * sizeof(uint4);
      // the XS instantiation also carries the staging loop for short gather lists (asd_runs.cuh): layouts with few neighbours take it
      // whether or not they have DM / BQ tables (ASD_SHORT_STAGING=0: only layouts with such tables)
      static const bool short_env = !(std::getenv("ASD_SHORT_STAGING") && atoi(std::getenv("ASD_SHORT_STAGING")) == 0);
      const bool xs = L.t.dm16 != nullptr || L.t.bq16 != nullptr || (short_env && L.t.ucap <= 6 * 256);
      static const bool pdl_env = !(std::getenv("ASD_PDL") && atoi(std::getenv("ASD_PDL")) == 0);
      // (slabs: dependent launches measured no gain with the side-stream exchange and a loss after the fused boundary launch,
      // 1.97 against 1.87 ms per step on a 512 x 512 x 32 slab: off)
      const bool pdl = pdl_env && !EDGE && !e->slab.on && big_grid;
#define ASD_LAUNCH_RUNS(NWV, XSV, LEANV, MMV)                                                                                                \
      do {                                                                                                                         \
         allow_smem(llg_runs_kernel<SOLVER, STAGE, NWV, EDGE, MSUM, XSV, LEANV, MMV>, smem);                                                   \
         if (pdl) {                                                                                                                \
            /* programmatic dependent launch: the CTAs of this stage become resident while the previous stage drains its last */  \
            /* wave, run the prologue that only touches tables and wait (griddepcontrol.wait) before the first spin is read */    \
            cudaLaunchConfig_t cfg;                                                                                                \
            memset(&cfg, 0, sizeof cfg);                                                                                           \
            cfg.gridDim = g; cfg.blockDim = dim3(NWV * 32, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;             \
            cudaLaunchAttribute at[1];                                                                                             \
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                         \
            at[0].val.programmaticStreamSerializationAllowed = 1;                                                                  \
            cfg.attrs = at; cfg.numAttrs = 1;                                                                                      \
            cudaLaunchKernelEx(&cfg, llg_runs_kernel<SOLVER, STAGE, NWV, EDGE, MSUM, XSV, LEANV, MMV>, L.t, p, ep, tr, e->cur.p, e->pred.p,    \
                               e->b2eff.p);                                                                                        \
         } else                                                                                                                    \
            llg_runs_kernel<SOLVER, STAGE, NWV, EDGE, MSUM, XSV, LEANV, MMV><<<g, NWV * 32, smem, e->stream>>>(L.t, p, ep, tr, e->cur.p, e->pred.p, e->b2eff.p); \
      } while (0)
      // plain Heisenberg system, uniform field and LLG parameters: the instantiation without the runtime checks of the general form
      static const bool lean_env = !(std::getenv("ASD_LEAN") && atoi(std::getenv("ASD_LEAN")) == 0);
      // (only together with the moment planes: without them the LEAN loop measured SLOWER than the general one, 0.487 against
      // 0.456 ms per step -- ptxas then issues the first spin loads of the staging loop after all fourteen index loads)
      const bool lean = lean_env && mm && !xs && NW == 8 && !p.per_site && p.mompar == 0 && L.t.btorque == nullptr && L.t.ext_uniform &&
                        L.t.zdm == 0 && L.t.zbq == 0 && !L.t.jtens;
      // XS layouts (DM / BQ tables, short lists): the general field terms with the lean integrator (uniform LLG parameters, no torque field)
      const bool ilean = lean_env && !EDGE && NW == 8 && !p.per_site && p.mompar == 0 && L.t.btorque == nullptr;
      if (NW == 8) {
         if (xs && ilean) ASD_LAUNCH_RUNS(8, true, (EDGE ? 0 : 3), false);
         else if (xs) ASD_LAUNCH_RUNS(8, true, 0, false);
         else if (mm && lean && L.t.do_aniso) ASD_LAUNCH_RUNS(8, false, 2, true);
         else if (mm && lean) ASD_LAUNCH_RUNS(8, false, 1, true);
         else if (mm) ASD_LAUNCH_RUNS(8, false, 0, true);
         else ASD_LAUNCH_RUNS(8, false, 0, false);
      }
      else if (NW == 4) ASD_LAUNCH_RUNS(4, false, 0, false);
      else ASD_LAUNCH_RUNS(2, false, 0, false);
#undef ASD_LAUNCH_RUNS
   } else if (L.t.staged && !fr) {
      const size_t smem = L.smem_bytes + (size_t)3 * L.t.ucap * sizeof(double);
      static const bool lean_env2 = !(std::getenv("ASD_LEAN") && atoi(std::getenv("ASD_LEAN")) == 0);
      const bool ilean2 = lean_env2 && !EDGE && !p.per_site && p.mompar == 0 && L.t.btorque == nullptr;
      if (L.reduced && ilean2) {
         // the lean integrator (asd_runs.cuh) in the one-atom-per-thread staged kernel of layouts without the run regularity
         allow_smem(llg_stage_kernel<SOLVER, STAGE, true, true, false, MSUM, true>, smem);
         launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, true, true, false, MSUM, true>, g, b, smem, e->stream, L.t, p, ep, tr, e->cur.p, e->pred.p, e->b2eff.p);
      } else if (ilean2) {
         // ... and with one coupling row per atom (do_reduced N, random alloys)
         allow_smem(llg_stage_kernel<SOLVER, STAGE, false, true, false, MSUM, true>, smem);
         launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, false, true, false, MSUM, true>, g, b, smem, e->stream, L.t, p, ep, tr, e->cur.p, e->pred.p, e->b2eff.p);
      } else if (L.reduced) {
         allow_smem(llg_stage_kernel<SOLVER, STAGE, true, true, EDGE, MSUM>, smem);
         launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, true, true, EDGE, MSUM>, g, b, smem, e->stream, L.t, p, ep, tr, e->cur.p, e->pred.p, e->b2eff.p);
      } else {
         allow_smem(llg_stage_kernel<SOLVER, STAGE, false, true, EDGE, MSUM>, smem);
         launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, false, true, EDGE, MSUM>, g, b, smem, e->stream, L.t, p, ep, tr, e->cur.p, e->pred.p, e->b2eff.p);
      }
   } else if (L.reduced) launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, true, false, EDGE, MSUM>, g, b, L.smem_bytes, e->stream, L.t, p, ep, trd, e->cur.p, e->pred.p, e->b2eff.p);
   else launch_dep(pdl_s, llg_stage_kernel<SOLVER, STAGE, false, false, EDGE, MSUM>, g, b, (size_t)0, e->stream, L.t, p, ep, trd, e->cur.p, e->pred.p, e->b2eff.p);
   e->launches++;
}

template <int SOLVER, int STAGE, bool EDGE>
static void launch_stage_range(asd_engine* e, Layout& L, const LlgParams& p, const EdgeParams& ep, const TileRange& tr, int ntiles) {
   if (ntiles <= 0) return;
   // the per-tile moment sums ride on corrector launches only
   if (STAGE == 2 && p.msum_part != nullptr) launch_stage_range2<SOLVER, STAGE, EDGE, STAGE == 2>(e, L, p, ep, tr, ntiles);
   else launch_stage_range2<SOLVER, STAGE, EDGE, false>(e, L, p, ep, tr, ntiles);
}

// boundary / interior split of a slab in tiles of `ts` slots: tiles [0,a) and [b,ntile) hold the H boundary planes
static void slab_split(const asd_engine* e, int ts, int& a, int& b, int& ntile) {
   const LatticeDesc& d = e->lat;
   ntile = (d.Nown + ts - 1) / ts;
   const long layer = (long)d.NTX * d.NSY * d.SY * d.SZ * d.NA * d.P;   // slots per layer of super-bricks
   const long nl = (d.H + d.SZ * d.BZ - 1) / (d.SZ * d.BZ);             // layers that hold the H boundary planes
   long aa = (layer * nl + ts - 1) / ts, bb = (d.Nown - layer * nl) / ts;
   if (bb < aa) { aa = ntile; bb = ntile; }                             // thin slab: every tile is a boundary tile
   a = (int)aa; b = (int)bb;
}

static EdgeParams edge_params(asd_engine* e, int stage, unsigned long long epoch) {
   Slab& sb = e->slab;
   EdgeParams ep;
   memset(&ep, 0, sizeof ep);
   ep.hdst_lo = sb.hdst_lo.p; ep.hdst_hi = sb.hdst_hi.p;
   ep.peer_lo = (stage == 1) ? sb.peer_pred[0] : sb.peer_cur[0];
   ep.peer_hi = (stage == 1) ? sb.peer_pred[1] : sb.peer_cur[1];
   ep.peer_mlo = (stage == 1) ? sb.peer_mpred[0] : sb.peer_mcur[0];
   ep.peer_mhi = (stage == 1) ? sb.peer_mpred[1] : sb.peer_mcur[1];
   // this rank is the UPPER neighbour of its lower neighbour: it owns word [1] there, and word [0] above
   ep.flag_lo = e->lat.has_lo ? sb.peer_flags[0] + 1 : nullptr;
   ep.flag_hi = e->lat.has_hi ? sb.peer_flags[1] + 0 : nullptr;
   ep.epoch = epoch;
   ep.ctr = sb.ctr.p;
   return ep;
}

// one stage over the whole engine: plain launch, or (slab) wait for the halos -> boundary tiles with the fused halo
// push -> interior tiles, which overlap the NVLink stores of the boundary launch
// halo exchange of a moment-plane layout: emomM of the atoms in the boundary tiles [0, ta) and [tb, nt) that a ring neighbour mirrors,
// from this slab's planes P[M][3][Npad] into the neighbours' planes (peer stores over NVLink); the last CTA publishes the epoch
__global__ void __launch_bounds__(256)
halo_push_planes_kernel(int ta, int tb, int nt, int ts, int Nown, size_t Npad, const double* __restrict__ P, EdgeParams ep) {
   const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
   const int k = blockIdx.y;
   const long tq = q / ts;
   if (tq < (long)ta + (nt - tb)) {
      const long tile = tq < ta ? tq : (long)tb + (tq - ta);
      const long i = tile * ts + q % ts;
      if (i < Nown) {
         const int lo = __ldg(ep.hdst_lo + i), hi = __ldg(ep.hdst_hi + i);
         if (lo >= 0 || hi >= 0) {
            const double* __restrict__ src = P + (size_t)k * 3 * Npad + i;
            const double mx = src[0], my = src[Npad], mz = src[2 * Npad];
            if (lo >= 0) { double* __restrict__ d = ep.peer_mlo + (size_t)k * 3 * Npad + lo; d[0] = mx; d[Npad] = my; d[2 * Npad] = mz; }
            if (hi >= 0) { double* __restrict__ d = ep.peer_mhi + (size_t)k * 3 * Npad + hi; d[0] = mx; d[Npad] = my; d[2 * Npad] = mz; }
         }
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

// main stream waits for the last halo push of the side stream (before the buffer it read is rewritten / before anything else uses ep.ctr)
static void slab_join_push(asd_engine* e) {
   Slab& sb = e->slab;
   if (sb.push_pending) { cudaStreamWaitEvent(e->stream, sb.evP, 0); sb.push_pending = false; }
}

template <int SOLVER, int STAGE>
static void launch_stage(asd_engine* e, Layout& L, const LlgParams& p) {
   Slab& sb = e->slab;
   const int ntile = (L.t.Nown + L.t.tile_slots - 1) / L.t.tile_slots;
   EdgeParams none;
   memset(&none, 0, sizeof none);
   if (!sb.on) {
      launch_stage_range<SOLVER, STAGE, false>(e, L, p, none, TileRange{0, ntile, 0}, ntile);
      return;
   }
   static const bool side_env = !(std::getenv("ASD_SLAB_SIDE") && atoi(std::getenv("ASD_SLAB_SIDE")) == 0);
   if (side_env && L.t.runs && L.t.mm && p.mm_cur != nullptr && sb.push_stream != nullptr) {
      // moment-plane layout: boundary tiles and interior tiles take the SAME kernels (no remote stores, no fences in them); the
      // boundary planes cross NVLink in a small launch of their own on a high-priority side stream while the interior tiles run
      slab_join_push(e);
      halo_wait_kernel<<<1, 1, 0, e->stream>>>(sb.flags.p, e->lat.has_lo, e->lat.has_hi, sb.epoch, sb.timeout_ticks, sb.err.p);
      e->launches++;
      int ta, tb, nt;
      slab_split(e, L.t.tile_slots, ta, tb, nt);
      const int nb = ta + (nt - tb);
      launch_stage_range<SOLVER, STAGE, false>(e, L, p, none, TileRange{0, ta, tb}, nb);
      cudaEventRecord(sb.evB, e->stream);
      launch_stage_range<SOLVER, STAGE, false>(e, L, p, none, TileRange{ta, tb - ta, 0}, tb - ta);
      cudaStreamWaitEvent(sb.push_stream, sb.evB, 0);
      const EdgeParams ep = edge_params(e, STAGE, sb.epoch + 1);
      const long nslot = (long)nb * L.t.tile_slots;
      halo_push_planes_kernel<<<dim3((unsigned)((nslot + 255) / 256), e->M), 256, 0, sb.push_stream>>>(
         ta, tb, nt, L.t.tile_slots, L.t.Nown, (size_t)L.Npad, (STAGE == 1) ? p.mm_pred : p.mm_cur, ep);
      e->launches++;
      cudaEventRecord(sb.evP, sb.push_stream);
      sb.push_pending = true;
      sb.epoch += 1;
      return;
   }
   halo_wait_kernel<<<1, 1, 0, e->stream>>>(sb.flags.p, e->lat.has_lo, e->lat.has_hi, sb.epoch, sb.timeout_ticks, sb.err.p);
   e->launches++;
   const EdgeParams ep = edge_params(e, STAGE, sb.epoch + 1);
   int ta, tb, nt;
   slab_split(e, L.t.tile_slots, ta, tb, nt);
   launch_stage_range<SOLVER, STAGE, true>(e, L, p, ep, TileRange{0, ta, tb}, ta + (nt - tb));
   launch_stage_range<SOLVER, STAGE, false>(e, L, p, none, TileRange{ta, tb - ta, 0}, tb - ta);
   sb.epoch += 1;
}

// copies the boundary planes of `cur` into the neighbours' halos (after the state was set from outside)
static int slab_push_state(asd_engine* e) {
   Slab& sb = e->slab;
   if (!sb.on) return 0;
   if (!sb.connected) return fail(-11, "slab: asd_slab_connect_* must be called before the moments are set");
   Layout& L = e->sd;
   slab_join_push(e);
   const EdgeParams ep = edge_params(e, 2, sb.epoch + 1);
   halo_push_kernel<<<dim3((L.t.Nown + 255) / 256, e->M), 256, 0, e->stream>>>(L.t.Nown, e->M, (size_t)L.Npad, e->cur.p, ep);
   e->launches++;
   CU(cudaGetLastError());
   sb.epoch += 1;
   return 0;
}

// Small systems: the whole time loop in one launch, state resident in shared memory, one thread-block cluster per
// ensemble (llg_resident_kernel).  Applies when one ensemble's cur + pred (64 bytes per slot) and the staged couplings
// fit the shared memory of one SM and the serial work per thread does not outweigh the launches it saves: with
// apt = atoms per thread (1 up to 8 x 256 = 2048 atoms) and a dependent chain of ~(1.7 + 0.09 z) us per atom and stage
// (measured, profiles/README.md), against ~3.3 us saved per stage.  ASD_RESIDENT=0 switches it off, =1 forces it
// (A/B runs, tests of either path on small fixtures).
struct ResidentPlan { size_t smem; int nrank, nt, apt; };
static const int RESIDENT_UNAVAILABLE = 31415;   // launch_resident: the cluster cannot be scheduled, use the stage launches

static bool resident_applies(const asd_engine* e, const Layout& L, ResidentPlan& rp) {
   const char* env = std::getenv("ASD_RESIDENT");
   if (env && atoi(env) == 0) return false;
   if (e->slab.on || L.t.runs || L.t.jtens) return false;   // run-compressed layouts are big-system layouts (tests build small ones on purpose)
   const int nown = L.t.Nown > 0 ? L.t.Nown : L.Npad;
   if (L.Npad > 65535) return false;                        // 16-bit neighbour slots
   rp.nrank = std::min(8, std::max(1, (nown + 255) / 256));
   rp.nt = std::min(256, std::max(64, (((nown + rp.nrank - 1) / rp.nrank + 31) / 32) * 32));
   rp.apt = (nown + rp.nrank * rp.nt - 1) / (rp.nrank * rp.nt);
   const int zt = L.t.z + L.t.zdm + L.t.zbq;
   rp.smem = (size_t)(((L.t.sm_cp + L.t.sm_dm + L.t.sm_bq + 3) & ~3)) * sizeof(double) + (size_t)2 * L.Npad * sizeof(SpinVec) +
             (size_t)rp.apt * zt * rp.nt * sizeof(unsigned short);
   if (rp.smem > (size_t)220 * 1024) return false;
   const double chain = 1.7 + 0.03 * zt;                     // us per atom and stage (measured, profiles/README.md)
   return (env && atoi(env) == 1) || (rp.apt - 1) * chain < 3.3;
}

template <int SOLVER, bool REDUCED>
static int launch_resident2(asd_engine* e, const Tables& t, const LlgParams& p, const ResidentPlan& rp, long nsteps, long first_step) {
   allow_smem(llg_resident_kernel<SOLVER, REDUCED>, rp.smem);
   cudaLaunchConfig_t cfg;
   memset(&cfg, 0, sizeof cfg);
   cfg.gridDim = dim3((unsigned)(e->M * rp.nrank), 1, 1);
   cfg.blockDim = dim3((unsigned)rp.nt, 1, 1);
   cfg.dynamicSmemBytes = rp.smem;
   cfg.stream = e->stream;
   cudaLaunchAttribute at[1];
   at[0].id = cudaLaunchAttributeClusterDimension;
   at[0].val.clusterDim.x = (unsigned)rp.nrank; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
   cfg.attrs = at; cfg.numAttrs = 1;
   // can the device co-schedule one such cluster at all (shared memory per CTA x cluster size inside one GPC)?  If not,
   // the caller falls back to the stage launches instead of failing.
   int nclusters = 0;
   if (cudaOccupancyMaxActiveClusters(&nclusters, llg_resident_kernel<SOLVER, REDUCED>, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      return RESIDENT_UNAVAILABLE;
   }
   CU(cudaLaunchKernelEx(&cfg, llg_resident_kernel<SOLVER, REDUCED>, t, p, e->cur.p, e->pred.p, e->b2eff.p, nsteps,
                         (unsigned long long)first_step, rp.apt));
   e->launches++;
   return 0;
}

template <int SOLVER>
static int launch_resident(asd_engine* e, Layout& L, const LlgParams& p, const ResidentPlan& rp, long nsteps, long first_step) {
   Tables t = L.t;
   t.nl4 = nullptr; t.cp4 = nullptr; t.cpl_param = 0; t.staged = 0; t.runs = 0;   // the plain j = 1..n loop of site_field
   if (t.Nown <= 0) t.Nown = L.Npad;
   return L.reduced ? launch_resident2<SOLVER, true>(e, t, p, rp, nsteps, first_step)
                    : launch_resident2<SOLVER, false>(e, t, p, rp, nsteps, first_step);
}

// moment planes of the MM run kernels: allocated on first use, rebuilt from `cur` whenever something else wrote the spins
__global__ void moment_planes_kernel(size_t Npad, int M, const SpinVec* __restrict__ S, double* __restrict__ P) {
   const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   const int k = blockIdx.y;
   if (i < Npad) {
      const SpinVec v = S[(size_t)k * Npad + i];
      double* __restrict__ W = P + (size_t)k * 3 * Npad + i;
      W[0] = v.x * v.m; W[Npad] = v.y * v.m; W[2 * Npad] = v.z * v.m;
   }
}

static int mm_prepare(asd_engine* e, Layout& L) {
   if (!(L.t.runs && L.t.mm) || !e->frozen.empty()) { e->mm_valid = false; return 0; }   // (fixed-moment runs take the direct kernel)
   const size_t n = (size_t)3 * L.Npad * e->M;
   int r;
   if (e->mm_cur.n < n || e->mm_pred.n < n) {
      if ((r = e->mm_cur.alloc(n))) return r;
      if ((r = e->mm_pred.alloc(n))) return r;
      CU(cudaMemsetAsync(e->mm_pred.p, 0, n * sizeof(double), e->stream));
      e->mm_valid = false;
   }
   static const bool always = std::getenv("ASD_MM_ALWAYS") && atoi(std::getenv("ASD_MM_ALWAYS")) != 0;
   if (!e->mm_valid || always) {
      if (e->slab.on && e->slab.connected) {
         // the halo slots of `cur` are the neighbours' to fill: their last exchange must have landed before the planes are derived
         Slab& sb = e->slab;
         halo_wait_kernel<<<1, 1, 0, e->stream>>>(sb.flags.p, e->lat.has_lo, e->lat.has_hi, sb.epoch, sb.timeout_ticks, sb.err.p);
         e->launches++;
      }
      moment_planes_kernel<<<dim3((unsigned)((L.Npad + 255) / 256), e->M), 256, 0, e->stream>>>((size_t)L.Npad, e->M, e->cur.p, e->mm_cur.p);
      e->launches++;
      CU(cudaGetLastError());
      e->mm_valid = true;
   }
   return 0;
}

static int sd_steps(asd_engine* e, long nsteps, long first_step) {
   int r = ensure_layout(e, 1);
   if (r) return r;
   Layout& L = e->sd;
   if (e->SDEalgh != 1 && e->SDEalgh != 5) return fail(-7, "SDEalgh %d is not on this path (1 = midpoint, 5 = Depondt)", e->SDEalgh);
   if (e->SDEalgh == 5 && e->b2eff.n < (size_t)3 * L.Npad * e->M) { if ((r = e->b2eff.alloc((size_t)3 * L.Npad * e->M))) return r; }
   if ((r = mm_prepare(e, L))) return r;
   LlgParams p;
   if ((r = fill_llg(e, L, p, 0))) return r;
   if (e->slab.on && !e->slab.connected) return fail(-11, "slab: not connected to the ring neighbours");
   const int ntile = (L.t.Nown + L.t.tile_slots - 1) / L.t.tile_slots;
   ResidentPlan rp;
   if (nsteps > 0 && resident_applies(e, L, rp)) {
      e->msum_fresh = false;
      e->mm_valid = false;
      const int rr = e->SDEalgh == 1 ? launch_resident<1>(e, L, p, rp, nsteps, first_step) : launch_resident<5>(e, L, p, rp, nsteps, first_step);
      if (rr != RESIDENT_UNAVAILABLE) return rr;
   }
   if (nsteps > 0) {
      if ((r = e->msum_part.alloc((size_t)e->M * ntile * 4))) return r;
      e->msum_fresh = false;
   }
   for (long s = 0; s < nsteps; s++) {
      p.step = (unsigned long long)(first_step + s);
      {
         const long long q = (long long)p.step - e->tf_first;
         const bool on = e->tf_n > 0 && q >= 0 && q < e->tf_n;
         for (int a = 0; a < 3; a++) p.tf[a] = on ? e->h_tfield[(size_t)3 * q + a] : 0.0;
      }
      const bool last = (s == nsteps - 1);
      // the last corrector launch also leaves the per-tile sums of emomM (not on the fixed-moment path, whose tiles differ)
      p.msum_part = (last && p.frozen == nullptr) ? e->msum_part.p : nullptr;
      p.msum_ntile = ntile;
      if (e->SDEalgh == 1) { launch_stage<1, 1>(e, L, p); launch_stage<1, 2>(e, L, p); }
      else { launch_stage<5, 1>(e, L, p); launch_stage<5, 2>(e, L, p); }
   }
   if (nsteps > 0) { e->msum_fresh = (p.frozen == nullptr); e->msum_ntile = ntile; }
   CU(cudaGetLastError());
   // a slab on the moment planes exchanged emomM only: bring the SPINS of the neighbours' halos up to date once per call (field /
   // energy evaluation and the Monte Carlo tile kernels read them)
   if (nsteps > 0 && e->slab.on && p.mm_cur != nullptr && (r = slab_push_state(e))) return r;
   return 0;
}

static int pinned_red(asd_engine* e, size_t n) {
   if (e->h_red && e->h_red_n >= n) return 0;
   if (e->h_red) cudaFreeHost(e->h_red);
   e->h_red = nullptr; e->h_red_n = 0;
   CU(cudaMallocHost((void**)&e->h_red, n * sizeof(double)));
   e->h_red_n = n;
   return 0;
}

static int measure(asd_engine* e, Layout& L, double* msum, double* energy) {
   int r;
   if ((r = pinned_red(e, (size_t)e->M * 4))) return r;
   if (!energy && msum && e->msum_fresh && e->state_layout == 1 && &L == &e->sd) {
      // the corrector launch of the last step already reduced every tile: add the partials
      if ((r = e->red.alloc((size_t)e->M * 4))) return r;
      moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(e->msum_ntile, e->msum_part.p, e->red.p);
      e->launches++;
      CU(cudaGetLastError());
      double* h = e->h_red;
      CU(cudaMemcpyAsync(h, e->red.p, (size_t)e->M * 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      for (int k = 0; k < e->M; k++) for (int a = 0; a < 3; a++) msum[3 * k + a] = h[4 * k + a];
      return slab_check(e);
   }
   const int nblk = std::min(1184, (L.Npad + 255) / 256);
   if ((r = e->part.alloc((size_t)e->M * nblk * 4))) return r;
   if ((r = e->red.alloc((size_t)e->M * 4))) return r;
   dim3 g, b;
   double* es = nullptr;
   if (energy) {
      if ((r = e->esite.alloc((size_t)e->M * L.Npad))) return r;
      launch_cfg(L.Npad, e->M, g, b);
      if (L.reduced) field_kernel<true><<<g, b, L.smem_bytes, e->stream>>>(L.t, e->cur.p, nullptr, nullptr, nullptr, e->esite.p);
      else field_kernel<false><<<g, b, 0, e->stream>>>(L.t, e->cur.p, nullptr, nullptr, nullptr, e->esite.p);
      e->launches++;
      es = e->esite.p;
   }
   moment_partial_kernel<<<dim3(nblk, e->M), 256, 0, e->stream>>>(L.Npad, L.d_orig.p, e->cur.p, es, e->part.p);
   moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(nblk, e->part.p, e->red.p);
   e->launches += 2;
   CU(cudaGetLastError());
   double* h = e->h_red;
   CU(cudaMemcpyAsync(h, e->red.p, (size_t)e->M * 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   for (int k = 0; k < e->M; k++) {
      if (msum) for (int a = 0; a < 3; a++) msum[3 * k + a] = h[4 * k + a];
      if (energy) energy[k] = h[4 * k + 3] * e->mub / e->mry;
   }
   return slab_check(e);
}

// Monte Carlo on the lattice (brick) layout with a periodic colouring: the path of a slab-decomposed supercell
// (colours are global, one halo exchange per colour).  Per colour: wait for the halos -> boundary tiles
// with the fused halo push -> interior tiles.  Also selectable on an undecomposed device-built lattice
// (ASD_MC_TILES=1), where it runs the very same Markov chain as any slab decomposition of that supercell.
static int mc_sweeps_tiles(asd_engine* e, McParams& p, long nsweeps, long first_sweep) {
   int r = ensure_layout(e, 1);
   if (r) return r;
   if (e->lat_ncol == 0 && (r = lattice_colours(e))) return r;
   Layout& L = e->sd;
   Slab& sb = e->slab;
   if (sb.on && !sb.connected) return fail(-11, "slab: not connected to the ring neighbours");
   const int ntile = (L.t.Nown + 255) / 256;
   EdgeParams none;
   memset(&none, 0, sizeof none);
   auto launch = [&](bool edge, const EdgeParams& ep, const TileRange& tr, int ntiles) {
      if (ntiles <= 0) return;
      const dim3 g(ntiles, e->M), b(256);
      if (L.reduced) {
         if (edge) mc_tile_kernel<true, true><<<g, b, L.smem_bytes, e->stream>>>(L.t, p, ep, tr, e->lat_col.p, e->cur.p);
         else mc_tile_kernel<true, false><<<g, b, L.smem_bytes, e->stream>>>(L.t, p, ep, tr, e->lat_col.p, e->cur.p);
      } else {
         if (edge) mc_tile_kernel<false, true><<<g, b, 0, e->stream>>>(L.t, p, ep, tr, e->lat_col.p, e->cur.p);
         else mc_tile_kernel<false, false><<<g, b, 0, e->stream>>>(L.t, p, ep, tr, e->lat_col.p, e->cur.p);
      }
      e->launches++;
   };
   for (long s = 0; s < nsweeps; s++)
      for (int c = 0; c < e->lat_ncol; c++) {
         p.sweep = (unsigned long long)(first_sweep + s);
         p.colour = c;
         if (!sb.on) { launch(false, none, TileRange{0, ntile, 0}, ntile); continue; }
         halo_wait_kernel<<<1, 1, 0, e->stream>>>(sb.flags.p, e->lat.has_lo, e->lat.has_hi, sb.epoch, sb.timeout_ticks, sb.err.p);
         e->launches++;
         const EdgeParams ep = edge_params(e, 2, sb.epoch + 1);
         int ta, tb, nt;
         slab_split(e, 256, ta, tb, nt);
         launch(true, ep, TileRange{0, ta, tb}, ta + (nt - tb));
         launch(false, none, TileRange{ta, tb - ta, 0}, tb - ta);
         sb.epoch += 1;
      }
   CU(cudaGetLastError());
   return 0;
}

static bool mc_on_tiles(const asd_engine* e) {
   if (!e->lattice_built) return false;
   if (e->slab.on) return true;
   if (e->mc_layout == 2) return false;      // block sweep requested but not applicable: colour-major layout
   if (e->mc_layout >= 0) return e->mc_layout == 1;
   const char* env = std::getenv("ASD_MC_TILES");
   return env && atoi(env) != 0;
}

static int mc_sweeps(asd_engine* e, char mode, long nsweeps, long first_sweep, double temperature, double temprescale,
                     const double* extfield) {
   if (mode != 'M' && mode != 'H') return fail(-8, "MC mode '%c' is not on this path ('M' Metropolis, 'H' heat bath)", mode);
   if (e->jtensor) return fail(-8, "Monte Carlo with tensorial exchange (do_jtensor 1) is not on this path");
   McParams p;
   memset(&p, 0, sizeof p);
   p.mode = mode; p.temperature = temperature; p.temprescale = temprescale; p.k_bolt = e->k_bolt; p.mub = e->mub;
   for (int a = 0; a < 3; a++) p.extfield[a] = extfield ? extfield[a] : 0.0;
   p.seed = e->seed ^ 0x5bd1e995u;
   // delta=(2.0/25.0)*(k_bolt*temperature/mub)**(0.20_dblprec) (montecarlo.f90:142): 2.0/25.0 is a default-real constant
   // expression in the Fortran, i.e. the single-precision 0.08; ignores temprescale like the reference
   p.delta = (double)(2.0f / 25.0f) * std::pow(e->k_bolt * temperature / e->mub, 0.20);
   e->msum_fresh = false;
   e->mm_valid = false;
   if (mc_block_candidate(e)) {
      int rb = mc_block_prepare(e);
      if (rb) return rb;
      if (e->mcb.on) return mc_sweeps_block(e, p, nsweeps, first_sweep);
   }
   if (mc_on_tiles(e)) return mc_sweeps_tiles(e, p, nsweeps, first_sweep);
   int r = ensure_layout(e, 2);
   if (r) return r;
   Layout& L = e->mc;
   const int ncol = (int)L.colour_first.size();
   // small systems: all colours of all sweeps in one launch, one CTA per ensemble, state in shared memory
   // (mc_resident_kernel; ASD_RESIDENT=0 keeps the colour launches)
   {
      const char* env = std::getenv("ASD_RESIDENT");
      long mx = 0;
      for (int c = 0; c < ncol; c++) mx = std::max(mx, (long)L.colour_count[c]);
      const size_t smem = (size_t)(((L.t.sm_cp + L.t.sm_dm + L.t.sm_bq + 3) & ~3)) * sizeof(double) + (size_t)L.Npad * sizeof(SpinVec);
      // a colour class of up to 512 atoms is one round of the CTA; more rounds only pay while they stay cheaper than launches
      const bool fits = smem <= (size_t)200 * 1024 && mx <= 1024 && nsweeps > 0 && ncol > 0;
      if (fits && !(env && atoi(env) == 0)) {
         int r;
         if (L.d_classes.n != (size_t)ncol) {
            std::vector<int2> cl(ncol);
            for (int c = 0; c < ncol; c++) cl[c] = make_int2(L.colour_first[c], L.colour_count[c]);
            if ((r = L.d_classes.upload(cl, e->stream))) return r;
         }
         p.sweep = (unsigned long long)first_sweep;
         Tables t = L.t;
         t.nl4 = nullptr; t.cp4 = nullptr; t.cpl_param = 0;      // plain j = 1..n neighbour loop, couplings staged in shared memory
         const int nt = (int)std::min(512L, std::max(64L, ((mx + 31) / 32) * 32));
         if (L.reduced) {
            allow_smem(mc_resident_kernel<true>, smem);
            mc_resident_kernel<true><<<e->M, nt, smem, e->stream>>>(t, p, L.d_classes.p, ncol, (int)nsweeps, e->cur.p);
         } else {
            allow_smem(mc_resident_kernel<false>, smem);
            mc_resident_kernel<false><<<e->M, nt, smem, e->stream>>>(t, p, L.d_classes.p, ncol, (int)nsweeps, e->cur.p);
         }
         e->launches++;
         CU(cudaGetLastError());
         return 0;
      }
   }
   // launch-bound regime (many small colour classes): one cooperative launch for all colours of all sweeps
   {
      long mx = 0;
      for (int c = 0; c < ncol; c++) mx = std::max(mx, (long)L.colour_count[c] * e->M);
      const char* env = std::getenv("ASD_MC_PERSISTENT");
      int sms = 148, coop = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
      bool use = L.t.nlrow && (L.reduced || L.t.cprow) && coop && ncol >= 12 && mx * 8 <= (long)sms * 2048 && L.t.z >= 32 && nsweeps > 0;
      if (env) use = use && atoi(env) != 0;
      if (use) {
         int r;
         if (L.d_classes.n != (size_t)ncol) {
            std::vector<int2> cl(ncol);
            for (int c = 0; c < ncol; c++) cl[c] = make_int2(L.colour_first[c], L.colour_count[c]);
            if ((r = L.d_classes.upload(cl, e->stream))) return r;
         }
         int lpa = 8;
         while (lpa < 32 && mx * lpa < (long)sms * 1024 && 4 * lpa <= L.t.z) lpa *= 2;
         const char* le = std::getenv("ASD_MC_LPA");
         if (le && (atoi(le) == 8 || atoi(le) == 16 || atoi(le) == 32)) lpa = atoi(le);
         p.sweep = (unsigned long long)first_sweep;
         const int2* cls = L.d_classes.p;
         int nc = ncol, ns = (int)nsweeps;
         SpinVec* curp = e->cur.p;
         void* args[] = {(void*)&L.t, (void*)&p, (void*)&cls, (void*)&nc, (void*)&ns, (void*)&curp};
         const void* fn = L.reduced ? (lpa == 8 ? (const void*)mc_sweeps_persistent_kernel<8, true> : lpa == 16 ? (const void*)mc_sweeps_persistent_kernel<16, true>
                                                                                                            : (const void*)mc_sweeps_persistent_kernel<32, true>)
                                    : (lpa == 8 ? (const void*)mc_sweeps_persistent_kernel<8, false> : lpa == 16 ? (const void*)mc_sweeps_persistent_kernel<16, false>
                                                                                                             : (const void*)mc_sweeps_persistent_kernel<32, false>);
         int per_sm = 0;
         CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, L.smem_bytes));
         if (per_sm > 0) {
            const long want = (mx * lpa + 255) / 256;
            const int grid = (int)std::max(1L, std::min((long)per_sm * sms, want));
            CU(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, L.smem_bytes, e->stream));
            e->launches++;
            CU(cudaGetLastError());
            return 0;
         }
      }
   }
   for (long s = 0; s < nsweeps; s++)
      for (int c = 0; c < ncol; c++) {
         p.sweep = (unsigned long long)(first_sweep + s);
         p.first = L.colour_first[c]; p.count = L.colour_count[c];
         if (p.count == 0) continue;
         dim3 g((p.count + 255) / 256, e->M), b(256);
         // small classes: LPA lanes per update so that a launch still fills the GPU (ASD_MC_LPA forces a width)
         int lpa = 1;
         if (L.t.nlrow && (L.reduced || L.t.cprow)) {
            const long attempts = (long)p.count * e->M;
            while (lpa < 32 && attempts * lpa < 300000L && 4 * lpa <= L.t.z) lpa *= 2;
            const char* env = std::getenv("ASD_MC_LPA");
            if (env) lpa = atoi(env);
         }
         if (lpa > 1) {
            const dim3 gc((unsigned)(((long)p.count * lpa + 255) / 256), e->M);
#define ASD_COOP(W) do { if (L.reduced) mc_colour_coop_kernel<W, true><<<gc, b, L.smem_bytes, e->stream>>>(L.t, p, e->cur.p); \
                        else mc_colour_coop_kernel<W, false><<<gc, b, 0, e->stream>>>(L.t, p, e->cur.p); } while (0)
            switch (lpa) {
               case 2: ASD_COOP(2); break;
               case 4: ASD_COOP(4); break;
               case 8: ASD_COOP(8); break;
               case 16: ASD_COOP(16); break;
               default: ASD_COOP(32); break;
            }
#undef ASD_COOP
         } else if (L.reduced) mc_colour_kernel<true><<<g, b, L.smem_bytes, e->stream>>>(L.t, p, e->cur.p, nullptr);
         else mc_colour_kernel<false><<<g, b, 0, e->stream>>>(L.t, p, e->cur.p, nullptr);
         e->launches++;
      }
   CU(cudaGetLastError());
   return 0;
}

// buffers a slab shares with its ring neighbours + the boundary / interior tile split
static int slab_commit(asd_engine* e) {
   Slab& sb = e->slab;
   Layout& L = e->sd;
   const LatticeDesc& d = e->lat;
   int r;
   if ((r = e->cur.alloc((size_t)L.Npad * e->M))) return r;
   if ((r = e->pred.alloc((size_t)L.Npad * e->M))) return r;
   // the moment planes are part of what the ring neighbours map (exported with cur / pred), whether or not this layout uses them
   if ((r = e->mm_cur.alloc((size_t)3 * L.Npad * e->M))) return r;
   if ((r = e->mm_pred.alloc((size_t)3 * L.Npad * e->M))) return r;
   CU(cudaMemset(e->mm_cur.p, 0, (size_t)3 * L.Npad * e->M * sizeof(double)));
   CU(cudaMemset(e->mm_pred.p, 0, (size_t)3 * L.Npad * e->M * sizeof(double)));
   e->mm_valid = false;
   if (sb.push_stream == nullptr) {
      int lo_pri = 0, hi_pri = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
      CU(cudaStreamCreateWithPriority(&sb.push_stream, cudaStreamNonBlocking, hi_pri));
      CU(cudaEventCreateWithFlags(&sb.evB, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&sb.evP, cudaEventDisableTiming));
   }
   sb.push_pending = false;
   if ((r = sb.flags.alloc(2))) return r;
   if ((r = sb.ctr.alloc(1))) return r;
   if ((r = sb.err.alloc(1))) return r;
   CU(cudaMemset(e->cur.p, 0, (size_t)L.Npad * e->M * sizeof(SpinVec)));
   CU(cudaMemset(e->pred.p, 0, (size_t)L.Npad * e->M * sizeof(SpinVec)));
   CU(cudaMemset(sb.flags.p, 0, 2 * sizeof(unsigned long long)));
   CU(cudaMemset(sb.ctr.p, 0, sizeof(unsigned int)));
   CU(cudaMemset(sb.err.p, 0, sizeof(int)));
   sb.epoch = 0;
   (void)d;
   return 0;
}

// ================================================================================================
// C ABI -- explicit API
// ================================================================================================
extern "C" {

const char* asd_last_error(void) { return g_err.c_str(); }

int asd_device_count(void) {
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

int asd_create(asd_engine** out, int device) {
   if (!out) return fail(-1, "null output pointer");
   *out = nullptr;
   int n = asd_device_count();
   if (n <= 0) return fail(-9, "no CUDA device available: this library has no CPU fallback");
   if (device < 0) { CU(cudaGetDevice(&device)); }
   if (device >= n) return fail(-9, "device %d requested but only %d present", device, n);
   CU(cudaSetDevice(device));
   asd_engine* e = new asd_engine();
   e->device = device;
   CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
   *out = e;
   return 0;
}

void asd_destroy(asd_engine* e) {
   if (!e) return;
   cudaSetDevice(e->device);
   if (e->stream) { cudaStreamSynchronize(e->stream); }
   if (e->slab.push_stream) { cudaStreamSynchronize(e->slab.push_stream); cudaStreamDestroy(e->slab.push_stream); cudaEventDestroy(e->slab.evB); cudaEventDestroy(e->slab.evP); }
   for (int q = 0; q < e->slab.n_opened; q++) cudaIpcCloseMemHandle(e->slab.opened[q]);
   if (e->h_red) cudaFreeHost(e->h_red);
   cudaStream_t s = e->stream;
   delete e;
   if (s) cudaStreamDestroy(s);
}

int asd_set_constants(asd_engine* e, double gamma, double k_bolt, double mub, double mry) {
   e->gamma = gamma; e->k_bolt = k_bolt; e->mub = mub; e->mry = mry;
   return 0;
}

int asd_set_system(asd_engine* e, int Natom, int Mensemble, int nHam, const int* aHam) {
   if (Natom <= 0 || Mensemble <= 0 || nHam <= 0 || nHam > Natom) return fail(-1, "bad sizes Natom=%d Mensemble=%d nHam=%d", Natom, Mensemble, nHam);
   e->N = Natom; e->M = Mensemble; e->NH = nHam;
   e->aHam.resize(Natom);
   for (int i = 0; i < Natom; i++) {
      e->aHam[i] = aHam ? aHam[i] : i + 1;
      if (e->aHam[i] < 1 || e->aHam[i] > nHam) return fail(-1, "aHam(%d)=%d outside 1..nHam", i + 1, e->aHam[i]);
   }
   e->landeg.assign(Natom, 1.0); e->lambda.assign(Natom, 0.05); e->temp.assign(Natom, 0.0);
   e->llg_uniform = true; e->llg_thermal = false;
   e->committed = false; e->sd_built = e->mc_built = false; e->state_layout = 0;
   return 0;
}

static int set_table(asd_engine* e, HostTable& T, int z, int ncomp, const int* list, const int* lsize, const double* coup) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (z <= 0 || !list || !lsize || !coup) return fail(-1, "bad table arguments");
   T.z = z; T.ncomp = ncomp;
   T.list.assign(list, list + (size_t)z * e->N);
   T.lsize.assign(lsize, lsize + e->NH);
   for (int h = 0; h < e->NH; h++) if (T.lsize[h] < 0 || T.lsize[h] > z) return fail(-1, "listsize(%d)=%d outside 0..%d", h + 1, T.lsize[h], z);
   T.coup.assign(coup, coup + (size_t)ncomp * z * e->NH);
   e->committed = false;
   return 0;
}
int asd_set_exchange(asd_engine* e, int z, const int* nlist, const int* nlistsize, const double* ncoup) {
   e->jtensor = false;
   return set_table(e, e->ex, z, 1, nlist, nlistsize, ncoup);
}
int asd_set_jtensor(asd_engine* e, int z, const int* nlist, const int* nlistsize, const double* j_tens) {
   e->jtensor = true;
   return set_table(e, e->ex, z, 9, nlist, nlistsize, j_tens);
}
int asd_set_dm(asd_engine* e, int z, const int* dmlist, const int* dmlistsize, const double* dm_vect) { return set_table(e, e->dm, z, 3, dmlist, dmlistsize, dm_vect); }
int asd_set_bq(asd_engine* e, int z, const int* bqlist, const int* bqlistsize, const double* j_bq) { return set_table(e, e->bq, z, 1, bqlist, bqlistsize, j_bq); }

int asd_set_lattice_hint(asd_engine* e, int NA, int N1, int N2, int N3, const char* bc3) {
   if (NA < 1 || N1 < 1 || N2 < 1 || N3 < 1 || !bc3) return fail(-1, "bad lattice hint");
   e->hint.on = 1; e->hint.NA = NA; e->hint.N1 = N1; e->hint.N2 = N2; e->hint.N3 = N3;
   for (int a = 0; a < 3; a++) e->hint.periodic[a] = (bc3[a] == 'P' || bc3[a] == 'p') ? 1 : 0;
   e->committed = false;
   return 0;
}

int asd_set_anisotropy(asd_engine* e, const int* taniso, const double* eaniso, const double* kaniso, const double* sb) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   e->taniso.assign(taniso, taniso + e->N);
   e->eaniso.assign(eaniso, eaniso + 3 * (size_t)e->N);
   e->kaniso.assign(kaniso, kaniso + 2 * (size_t)e->N);
   e->sb.assign(sb, sb + e->N);
   e->have_aniso = true; e->committed = false;
   return 0;
}

int asd_set_external_field(asd_engine* e, const double* f) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (f) e->ext.assign(f, f + 3 * (size_t)e->N * e->M); else e->ext.clear();
   if (e->committed) {
      CU(cudaSetDevice(e->device));
      CU(cudaStreamSynchronize(e->stream));
      int r = apply_external_field(e, e->sd);
      if (r) return r;
      if (e->mc_built && (r = apply_external_field(e, e->mc))) return r;
   }
   return 0;
}
int asd_set_torque(asd_engine* e, const double* f) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (f) e->btorque.assign(f, f + 3 * (size_t)e->N * e->M); else e->btorque.clear();
   e->committed = false;
   return 0;
}

int asd_set_time_field(asd_engine* e, long first_step, long nsteps, const double* tfield) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   CU(cudaSetDevice(e->device));
   if (!tfield || nsteps <= 0) { e->tf_n = 0; e->h_tfield.clear(); return 0; }
   if (nsteps > 100000000L) return fail(-1, "asd_set_time_field: schedule too long (%ld steps)", nsteps);
   const size_t n = (size_t)nsteps * 3;
   e->h_tfield.assign(tfield, tfield + n);
   int r = e->tfield.alloc(n);
   if (r) return r;
   CU(cudaMemcpyAsync(e->tfield.p, tfield, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   e->tf_first = first_step; e->tf_n = (int)nsteps;
   return 0;
}

int asd_set_llg(asd_engine* e, int SDEalgh, double delta_t, const double* Landeg, const double* lambda1_array,
                const double* Temp_array, double temprescale, int mompar, unsigned long long seed) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (SDEalgh != 1 && SDEalgh != 5) return fail(-7, "SDEalgh %d is not on this path (1 = midpoint, 5 = Depondt)", SDEalgh);
   e->SDEalgh = SDEalgh; e->delta_t = delta_t; e->temprescale = temprescale; e->mompar = mompar; e->seed = seed;
   if (Landeg) e->landeg.assign(Landeg, Landeg + e->N);
   if (lambda1_array) e->lambda.assign(lambda1_array, lambda1_array + e->N);
   if (Temp_array) e->temp.assign(Temp_array, Temp_array + e->N);
   e->sd.d_lambda.release(); e->sd.d_landeg.release(); e->sd.d_temp.release();
   auto uniform = [&](const std::vector<double>& v) { for (double x : v) if (x != v[0]) return false; return true; };
   e->llg_uniform = uniform(e->landeg) && uniform(e->lambda) && uniform(e->temp);
   e->llg_thermal = false;
   for (double x : e->temp) if (x > 0.0) e->llg_thermal = true;
   if (mompar != 0 && e->committed && e->state_layout == 1 && e->sd.d_mmom0.p == nullptr) {
      // moments were uploaded before mompar was switched on: re-stage through the host copy
      int r = stash_state_to_host(e);
      if (r) return r;
      e->state_layout = 0;
   }
   return 0;
}

int asd_set_evolving_atoms(asd_engine* e, int Nred, const int* red_atom_list) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   e->sd.d_frozen.release(); e->mc.d_frozen.release();
   e->frozen.clear();
   if (red_atom_list == nullptr || Nred <= 0 || Nred >= e->N) {
      if (red_atom_list != nullptr && Nred > e->N) return fail(-1, "asd_set_evolving_atoms: Nred = %d exceeds Natom", Nred);
      if (red_atom_list == nullptr || Nred <= 0) return 0;      // every atom evolves
   }
   if (e->slab.on) return fail(-11, "asd_set_evolving_atoms: not available on a slab");
   e->frozen.assign(e->N, 1);
   for (int q = 0; q < Nred; q++) {
      const int a = red_atom_list[q];
      if (a < 1 || a > e->N) { e->frozen.clear(); return fail(-1, "asd_set_evolving_atoms: atom %d outside 1..Natom", a); }
      e->frozen[a - 1] = 0;
   }
   bool any = false;
   for (unsigned char f : e->frozen) any = any || f;
   if (!any) e->frozen.clear();
   return 0;
}

int asd_set_moments(asd_engine* e, const double* emom, const double* mmom, const double* mmom0) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   const size_t NM = (size_t)e->N * e->M;
   if (e->committed && e->sd_built) {
      // tables are frozen: go straight from the caller's buffers to the device layout (no host-side copy)
      CU(cudaSetDevice(e->device));
      if (mmom0) e->h_mmom0.assign(mmom0, mmom0 + NM); else e->h_mmom0.clear();
      int r = upload_state_from(e, e->sd, emom, mmom, mmom0);
      if (r) return r;
      e->h_emom.clear(); e->h_mmom.clear();
      e->state_layout = 1;
      return 0;
   }
   e->h_emom.assign(emom, emom + 3 * NM);
   e->h_mmom.assign(mmom, mmom + NM);
   if (mmom0) e->h_mmom0.assign(mmom0, mmom0 + NM); else e->h_mmom0.clear();
   e->state_layout = 0;
   return 0;
}

int asd_get_moments(asd_engine* e, double* emom, double* emomM, double* mmom) {
   CU(cudaSetDevice(e->device));
   if (e->state_layout == 0) { int r = ensure_layout(e, 1); if (r) return r; }
   return download_state(e, e->state_layout == 1 ? e->sd : e->mc, emom, emomM, mmom);
}

int asd_commit(asd_engine* e) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   CU(cudaSetDevice(e->device));
   e->mm_valid = false;
   if (e->committed && !e->slab.on) { int r = stash_state_to_host(e); if (r) return r; }
   if (!e->lattice_built) {
      if (!e->ex.present()) return fail(-2, "no exchange table (asd_set_exchange / asd_build_lattice_table)");
      int r = build_layout(e, e->sd, false);
      if (r) return r;
   } else {
      if (e->sd.zs[0] == 0) return fail(-2, "no exchange table built");
      int r = finish_layout(e, e->sd);
      if (r) return r;
   }
   if (e->slab.on) {
      if (!e->lattice_built) return fail(-11, "slab decomposition needs the on-device lattice builder (asd_build_lattice_table)");
      int r = slab_commit(e);
      if (r) return r;
   }
   e->sd_built = true; e->mc_built = false;
   e->mcb.tried = false; e->mcb.on = false;
   e->committed = true;
   if (e->state_layout != 0) e->state_layout = 0;
   return 0;
}

int asd_effective_field(asd_engine* e, double* beff, double* beff1, double* beff2, double* energy) {
   CU(cudaSetDevice(e->device));
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   Layout& L = (e->state_layout == 2) ? e->mc : e->sd;
   const size_t NM3 = 3 * (size_t)e->N * e->M;
   DevBuf<double> d0, d1, d2;
   if (beff && (r = d0.alloc(NM3))) return r;
   if (beff1 && (r = d1.alloc(NM3))) return r;
   if (beff2 && (r = d2.alloc(NM3))) return r;
   if (energy && (r = e->esite.alloc((size_t)e->M * L.Npad))) return r;
   dim3 g, b;
   launch_cfg(L.Npad, e->M, g, b);
   if (L.reduced) field_kernel<true><<<g, b, L.smem_bytes, e->stream>>>(L.t, e->cur.p, d0.p, d1.p, d2.p, energy ? e->esite.p : nullptr);
   else field_kernel<false><<<g, b, 0, e->stream>>>(L.t, e->cur.p, d0.p, d1.p, d2.p, energy ? e->esite.p : nullptr);
   e->launches++;
   CU(cudaGetLastError());
   if (beff) CU(cudaMemcpyAsync(beff, d0.p, NM3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   if (beff1) CU(cudaMemcpyAsync(beff1, d1.p, NM3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   if (beff2) CU(cudaMemcpyAsync(beff2, d2.p, NM3 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   if (energy) {
      // reuse the reduction path without recomputing the field
      const int nblk = std::min(1184, (L.Npad + 255) / 256);
      if ((r = e->part.alloc((size_t)e->M * nblk * 4))) return r;
      if ((r = e->red.alloc((size_t)e->M * 4))) return r;
      moment_partial_kernel<<<dim3(nblk, e->M), 256, 0, e->stream>>>(L.Npad, L.d_orig.p, e->cur.p, e->esite.p, e->part.p);
      moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(nblk, e->part.p, e->red.p);
      e->launches += 2;
      std::vector<double> h((size_t)e->M * 4);
      CU(cudaMemcpyAsync(h.data(), e->red.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      for (int k = 0; k < e->M; k++) energy[k] = h[4 * k + 3] * e->mub / e->mry;
   }
   return 0;
}

int asd_sd_steps(asd_engine* e, long nsteps, long first_step) {
   CU(cudaSetDevice(e->device));
   return sd_steps(e, nsteps, first_step);
}

int asd_mc_sweeps(asd_engine* e, char mode, long nsweeps, long first_sweep, double temperature, double temprescale,
                  const double* extfield) {
   CU(cudaSetDevice(e->device));
   return mc_sweeps(e, mode, nsweeps, first_sweep, temperature, temprescale, extfield);
}

int asd_set_mc_layout(asd_engine* e, int layout) {
   if (layout < -1 || layout > 2) return fail(-1, "MC layout must be -1 (automatic), 0 (colour-major), 1 (lattice tiles) or 2 (block sweep)");
   if (layout == 2 && (!e->lattice_built || e->slab.on)) return fail(-2, "the block sweep needs an undecomposed device-built lattice (asd_build_lattice_table)");
   if (layout == 1 && !e->lattice_built) return fail(-2, "the lattice MC layout needs a device-built lattice (asd_build_lattice_table)");
   if (layout == 0 && e->slab.on) return fail(-5, "a slab runs Monte Carlo on the lattice layout only");
   e->mc_layout = layout;
   return 0;
}

int asd_mc_colouring(asd_engine* e, int* layout, int* ncolours, int* period3) {
   CU(cudaSetDevice(e->device));
   int r = mc_sweeps(e, 'M', 0, 1, 1.0, 1.0, nullptr);   // builds the layout / colouring, runs nothing
   if (r) return r;
   const bool block = mc_block_candidate(e) && e->mcb.on;
   const bool tiles = block || mc_on_tiles(e);
   if (layout) *layout = block ? 2 : tiles ? 1 : 0;
   if (ncolours) *ncolours = tiles ? e->lat_ncol : (int)e->mc.colour_first.size();
   if (period3) for (int a = 0; a < 3; a++) period3[a] = tiles ? e->lat_period[a] : 0;
   return 0;
}

int asd_get_mc_colours(asd_engine* e, int* colour) {
   CU(cudaSetDevice(e->device));
   int r = mc_sweeps(e, 'M', 0, 1, 1.0, 1.0, nullptr);
   if (r) return r;
   if ((mc_block_candidate(e) && e->mcb.on) || mc_on_tiles(e)) {
      Layout& L = e->sd;
      if ((r = host_orig(e, L))) return r;
      std::vector<unsigned char> c(L.t.Nown);
      CU(cudaMemcpy(c.data(), e->lat_col.p, c.size(), cudaMemcpyDeviceToHost));
      for (int s = 0; s < L.t.Nown; s++) if (L.orig[s] >= 0) colour[L.orig[s]] = c[s];
   } else {
      Layout& L = e->mc;
      for (size_t c = 0; c < L.colour_first.size(); c++)
         for (int s = L.colour_first[c]; s < L.colour_first[c] + L.colour_count[c]; s++)
            if (L.orig[s] >= 0) colour[L.orig[s]] = (int)c;
   }
   return 0;
}

int asd_get_mc_visit_order(asd_engine* e, int* order) {
   CU(cudaSetDevice(e->device));
   int r = mc_sweeps(e, 'M', 0, 1, 1.0, 1.0, nullptr);   // builds the layout / colouring, runs nothing
   if (r) return r;
   if (mc_block_candidate(e) && e->mcb.on) return mc_block_visit_order(e, order);
   size_t n = 0;
   if (mc_on_tiles(e)) {
      // one launch per colour over the lattice order: colour by colour, slots ascending
      Layout& L = e->sd;
      if ((r = host_orig(e, L))) return r;
      std::vector<unsigned char> c(L.t.Nown);
      CU(cudaMemcpy(c.data(), e->lat_col.p, c.size(), cudaMemcpyDeviceToHost));
      for (int col = 0; col < e->lat_ncol; col++)
         for (int s = 0; s < L.t.Nown; s++) if (L.orig[s] >= 0 && c[s] == col) order[n++] = L.orig[s] + 1;
   } else {
      Layout& L = e->mc;
      for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) order[n++] = L.orig[s] + 1;   // colour classes are slot ranges
   }
   if (n != (size_t)e->N) return fail(-5, "internal: visiting order covers %zu of %d atoms", n, e->N);
   return 0;
}

// the draws of one sweep, exactly as the update kernels compute them (test hook: a CPU restatement of mc_evolve replays a sweep with them)
__global__ void mc_draws_kernel(int N, int M, unsigned long long seed, unsigned long long sweep, unsigned int atom_offset,
                                unsigned int ens_offset, double* __restrict__ u, double* __restrict__ g) {
   const int o = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (o >= N) return;
   double uu[4], g0, g1, g2;
   uniform4(seed, (uint32_t)o + atom_offset, (uint32_t)k + ens_offset, sweep, 1u, uu);
   gauss3f(seed, (uint32_t)o + atom_offset, (uint32_t)k + ens_offset, sweep, 2u, g0, g1, g2);
   const size_t q = (size_t)o + (size_t)N * k;
   for (int a = 0; a < 4; a++) u[4 * q + a] = uu[a];
   g[3 * q] = g0; g[3 * q + 1] = g1; g[3 * q + 2] = g2;
}

int asd_debug_mc_draws(asd_engine* e, long sweep, double* u, double* g) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   CU(cudaSetDevice(e->device));
   const size_t NM = (size_t)e->N * e->M;
   DevBuf<double> du, dg;
   int r;
   if ((r = du.alloc(4 * NM))) return r;
   if ((r = dg.alloc(3 * NM))) return r;
   const unsigned int aoff = e->lattice_built ? e->sd.t.atom_offset : 0u;
   mc_draws_kernel<<<dim3((e->N + 255) / 256, e->M), 256, 0, e->stream>>>(e->N, e->M, e->seed ^ 0x5bd1e995u, (unsigned long long)sweep, aoff,
                                                                        e->ens_offset, du.p, dg.p);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaMemcpyAsync(u, du.p, 4 * NM * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaMemcpyAsync(g, dg.p, 3 * NM * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   return 0;
}

int asd_measure(asd_engine* e, double* msum, double* energy) {
   CU(cudaSetDevice(e->device));
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   return measure(e, e->state_layout == 2 ? e->mc : e->sd, msum, energy);
}

// Sampled run: the measurement-phase loop of sd_mphase (sd_driver.f90:517-849) without a host round trip per sample.  Every
// `sample_every` steps the per-ensemble sums of emomM (what prn_averages needs, prn_averages.f90:437-447) are reduced on the
// device into a sample ring; ONE device-to-host copy and ONE stream synchronisation end the call.  The corrector launch of a
// sampled step leaves the per-tile sums (no extra pass over the spins); small systems / fixed-moment runs use the stand-alone
// reduction.  The reference's CUDA loop instead copies the whole state to the host at every sampled step
// (cudaMdSimulation.cu:400-470, cudaMeasurement.cu:109-182).
static int sd_run(asd_engine* e, long nsteps, long first_step, long sample_every, double* msum, long* nsamples) {
   if (sample_every <= 0) sample_every = nsteps > 0 ? nsteps : 1;
   const long ns = nsteps / sample_every;
   int r = ensure_layout(e, 1);
   if (r) return r;
   Layout& L = e->sd;
   if (nsamples) *nsamples = ns;
   if (ns > 0) {
      if ((r = e->ring.alloc((size_t)ns * e->M * 4))) return r;
      if ((r = pinned_red(e, (size_t)ns * e->M * 4))) return r;
   }
   long done = 0;
   for (long q = 0; q < ns; q++) {
      if ((r = sd_steps(e, sample_every, first_step + done))) return r;
      done += sample_every;
      double* dst = e->ring.p + (size_t)q * e->M * 4;
      if (e->msum_fresh) {
         moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(e->msum_ntile, e->msum_part.p, dst);
         e->launches++;
      } else {
         const int nblk = std::min(1184, (L.Npad + 255) / 256);
         if ((r = e->part.alloc((size_t)e->M * nblk * 4))) return r;
         moment_partial_kernel<<<dim3(nblk, e->M), 256, 0, e->stream>>>(L.Npad, L.d_orig.p, e->cur.p, nullptr, e->part.p);
         moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(nblk, e->part.p, dst);
         e->launches += 2;
      }
   }
   if (done < nsteps && (r = sd_steps(e, nsteps - done, first_step + done))) return r;
   CU(cudaGetLastError());
   if (ns > 0 && msum) {
      CU(cudaMemcpyAsync(e->h_red, e->ring.p, (size_t)ns * e->M * 4 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      for (long q = 0; q < ns; q++)
         for (int k = 0; k < e->M; k++)
            for (int a = 0; a < 3; a++) msum[(size_t)3 * (k + (size_t)e->M * q) + a] = e->h_red[(size_t)4 * (k + (size_t)e->M * q) + a];
      return slab_check(e);
   }
   return 0;
}

int asd_sd_run(asd_engine* e, long nsteps, long first_step, long sample_every, double* msum, long* nsamples) {
   CU(cudaSetDevice(e->device));
   if (nsteps < 0) return fail(-1, "asd_sd_run: nsteps < 0");
   return sd_run(e, nsteps, first_step, sample_every, msum, nsamples);
}

int asd_time_sd_steps(asd_engine* e, long nsteps, long first_step, float* total_ms, float* stage_ms) {
   CU(cudaSetDevice(e->device));
   int r = ensure_layout(e, 1);
   if (r) return r;
   cudaEvent_t a, b;
   CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
   CU(cudaStreamSynchronize(e->stream));
   CU(cudaEventRecord(a, e->stream));
   r = sd_steps(e, nsteps, first_step);
   CU(cudaEventRecord(b, e->stream));
   CU(cudaEventSynchronize(b));
   if (total_ms) CU(cudaEventElapsedTime(total_ms, a, b));
   if (stage_ms && r == 0) {
      // time each stage kernel alone (one extra step, state advanced by it as well)
      Layout& L = e->sd;
      LlgParams p;
      if ((r = fill_llg(e, L, p, (unsigned long long)(first_step + nsteps)))) return r;
      e->msum_fresh = false;
      cudaEvent_t c;
      CU(cudaEventCreate(&c));
      CU(cudaEventRecord(a, e->stream));
      if (e->SDEalgh == 1) launch_stage<1, 1>(e, L, p); else launch_stage<5, 1>(e, L, p);
      CU(cudaEventRecord(b, e->stream));
      if (e->SDEalgh == 1) launch_stage<1, 2>(e, L, p); else launch_stage<5, 2>(e, L, p);
      CU(cudaEventRecord(c, e->stream));
      CU(cudaEventSynchronize(c));
      CU(cudaEventElapsedTime(&stage_ms[0], a, b));
      CU(cudaEventElapsedTime(&stage_ms[1], b, c));
      cudaEventDestroy(c);
      if (e->slab.on && p.mm_cur != nullptr && (r = slab_push_state(e))) return r;
   }
   cudaEventDestroy(a); cudaEventDestroy(b);
   return r;
}

#ifdef ASD_MC_PROF
extern "C" int asd_debug_mc_prof(unsigned long long* out8, int reset) {
   if (out8) CU(cudaMemcpyFromSymbol(out8, asd::g_mc_prof, 8 * sizeof(unsigned long long)));
   if (reset) { unsigned long long z[8] = {0}; CU(cudaMemcpyToSymbol(asd::g_mc_prof, z, sizeof z)); }
   return 0;
}
#endif

int asd_time_mc_sweeps(asd_engine* e, char mode, long nsweeps, double temperature, float* total_ms) {
   CU(cudaSetDevice(e->device));
   int r = mc_sweeps(e, mode, 0, 1, temperature, 1.0, nullptr);   // builds the layout / colouring outside the timed region
   if (r) return r;
   cudaEvent_t a, b;
   CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
   CU(cudaStreamSynchronize(e->stream));
   CU(cudaEventRecord(a, e->stream));
   r = mc_sweeps(e, mode, nsweeps, 1, temperature, 1.0, nullptr);
   CU(cudaEventRecord(b, e->stream));
   CU(cudaEventSynchronize(b));
   if (total_ms) CU(cudaEventElapsedTime(total_ms, a, b));
   cudaEventDestroy(a); cudaEventDestroy(b);
   return r;
}

int asd_energy_terms(asd_engine* e, double* terms) {
   CU(cudaSetDevice(e->device));
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   Layout& L = (e->state_layout == 2) ? e->mc : e->sd;
   const int R = 5, nblk = std::min(592, (L.Npad + 255) / 256);
   DevBuf<double> rows, part, out;
   if ((r = rows.alloc((size_t)e->M * R * L.Npad))) return r;
   if ((r = part.alloc((size_t)e->M * R * nblk))) return r;
   if ((r = out.alloc((size_t)e->M * R))) return r;
   dim3 g, b;
   launch_cfg(L.Npad, e->M, g, b);
   if (L.reduced) energy_terms_kernel<true><<<g, b, 0, e->stream>>>(L.t, e->cur.p, rows.p);
   else energy_terms_kernel<false><<<g, b, 0, e->stream>>>(L.t, e->cur.p, rows.p);
   reduce_rows_partial_kernel<<<dim3(nblk, R, e->M), 256, 0, e->stream>>>(L.Npad, R, rows.p, part.p);
   reduce_rows_final_kernel<<<e->M * R, 256, 0, e->stream>>>(nblk, part.p, out.p);
   e->launches += 3;
   CU(cudaGetLastError());
   std::vector<double> h((size_t)e->M * R);
   CU(cudaMemcpyAsync(h.data(), out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   const double fcinv = e->mub / e->mry;   // energy.f90: energies printed in mRy per atom
   for (size_t q = 0; q < h.size(); q++) terms[q] = h[q] * fcinv / e->N;
   return 0;
}

int asd_measure_sublattice(asd_engine* e, int NA, double* msum_na) {
   CU(cudaSetDevice(e->device));
   if (NA < 1 || e->N % NA) return fail(-1, "asd_measure_sublattice: NA = %d does not divide Natom", NA);
   if (e->slab.on) return fail(-11, "asd_measure_sublattice: not available on a slab (sum the slabs' asd_get_moments instead)");
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   Layout& L = (e->state_layout == 2) ? e->mc : e->sd;
   const int nblk = std::min(1184, (L.Npad + 255) / 256);
   DevBuf<double> out;
   if ((r = e->part.alloc((size_t)e->M * nblk * 4))) return r;
   if ((r = out.alloc((size_t)NA * e->M * 4))) return r;
   for (int c = 0; c < NA; c++) {
      moment_class_partial_kernel<<<dim3(nblk, e->M), 256, 0, e->stream>>>(L.Npad, L.d_orig.p, e->cur.p, NA, c, e->part.p);
      moment_final_kernel<<<e->M, 1024, 0, e->stream>>>(nblk, e->part.p, out.p + (size_t)c * e->M * 4);
      e->launches += 2;
   }
   CU(cudaGetLastError());
   std::vector<double> h((size_t)NA * e->M * 4);
   CU(cudaMemcpyAsync(h.data(), out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   for (int k = 0; k < e->M; k++)
      for (int c = 0; c < NA; c++)
         for (int a = 0; a < 3; a++) msum_na[a + 3 * (c + (size_t)NA * k)] = h[((size_t)c * e->M + k) * 4 + a];
   return 0;
}

int asd_set_triangulation(asd_engine* e, int nsimp, const int* simp) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (nsimp < 0) return fail(-1, "asd_set_triangulation: nsimp < 0");
   e->simp.resize((size_t)3 * nsimp);
   for (size_t q = 0; q < (size_t)3 * nsimp; q++) {
      if (simp[q] < 1 || simp[q] > e->N) return fail(-1, "asd_set_triangulation: atom %d outside 1..Natom", simp[q]);
      e->simp[q] = simp[q] - 1;
   }
   e->nsimp = nsimp; e->tri_layout = 0;
   return 0;
}

int asd_skyrmion_number(asd_engine* e, double* q) {
   CU(cudaSetDevice(e->device));
   if (e->nsimp <= 0) return fail(-2, "asd_skyrmion_number: asd_set_triangulation first");
   if (e->slab.on) return fail(-11, "asd_skyrmion_number: not available on a slab");
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   const int lay = (e->state_layout == 2) ? 2 : 1;
   Layout& L = (lay == 2) ? e->mc : e->sd;
   const int ns = e->nsimp;
   if (e->tri_layout != lay) {
      if ((r = host_orig(e, L))) return r;
      std::vector<int> tri((size_t)3 * ns);
      for (int t = 0; t < ns; t++)
         for (int c = 0; c < 3; c++) tri[(size_t)c * ns + t] = L.slot_of[e->simp[(size_t)3 * t + c]];   // simp(3, nsimp)
      if ((r = e->d_tri.upload(tri, e->stream))) return r;
      e->tri_layout = lay;
   }
   const int nblk = std::min(592, (ns + 255) / 256);
   DevBuf<double> rows, part, out;
   if ((r = rows.alloc((size_t)e->M * ns))) return r;
   if ((r = part.alloc((size_t)e->M * nblk))) return r;
   if ((r = out.alloc((size_t)e->M))) return r;
   skyrmion_tri_kernel<<<dim3((ns + 255) / 256, e->M), 256, 0, e->stream>>>(ns, (size_t)L.Npad, e->d_tri.p, e->cur.p, rows.p);
   reduce_rows_partial_kernel<<<dim3(nblk, 1, e->M), 256, 0, e->stream>>>(ns, 1, rows.p, part.p);
   reduce_rows_final_kernel<<<e->M, 256, 0, e->stream>>>(nblk, part.p, out.p);
   e->launches += 3;
   CU(cudaGetLastError());
   std::vector<double> h((size_t)e->M);
   CU(cudaMemcpyAsync(h.data(), out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   const double four_pi = 4.0 * 3.14159265358979323846;
   for (int k = 0; k < e->M; k++) q[k] = h[k] / four_pi;
   return 0;
}

int asd_get_atoms(asd_engine* e, int n, const int* atoms, double* out) {
   CU(cudaSetDevice(e->device));
   if (n <= 0) return 0;
   int r = ensure_layout(e, e->state_layout == 2 ? 2 : 1);
   if (r) return r;
   Layout& L = (e->state_layout == 2) ? e->mc : e->sd;
   if ((r = host_orig(e, L))) return r;
   std::vector<int> sl(n);
   for (int q = 0; q < n; q++) {
      if (atoms[q] < 1 || atoms[q] > e->N) return fail(-1, "atom %d outside 1..Natom", atoms[q]);
      sl[q] = L.slot_of[atoms[q] - 1];
   }
   DevBuf<int> d_sl;
   DevBuf<double> d_out;
   if ((r = d_sl.upload(sl, e->stream))) return r;
   if ((r = d_out.alloc((size_t)4 * n * e->M))) return r;
   gather_atoms_kernel<<<dim3((n + 127) / 128, e->M), 128, 0, e->stream>>>(n, e->M, (size_t)L.Npad, d_sl.p, e->cur.p, d_out.p);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaMemcpyAsync(out, d_out.p, (size_t)4 * n * e->M * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
   CU(cudaStreamSynchronize(e->stream));
   return 0;
}

int asd_set_ensemble_offset(asd_engine* e, unsigned int first_ensemble) {
   e->ens_offset = first_ensemble;
   e->sd.t.ens_offset = first_ensemble; e->mc.t.ens_offset = first_ensemble;
   return 0;
}

int asd_set_slab(asd_engine* e, int nslabs, int slab_index, int halo_planes) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   if (e->lattice_built) return fail(-2, "asd_set_slab must precede asd_build_lattice_table");
   if (nslabs < 1 || slab_index < 0 || slab_index >= nslabs || halo_planes < 1) return fail(-1, "bad slab arguments");
   e->slab.on = 1; e->slab.G = nslabs; e->slab.g = slab_index; e->slab.H = halo_planes;
   e->slab.connected = false;
   return 0;
}

int asd_slab_export(asd_engine* e, void* handles) {
   if (!e->slab.on || !e->committed) return fail(-11, "slab: asd_set_slab + asd_commit first");
   CU(cudaSetDevice(e->device));
   cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles;
   CU(cudaIpcGetMemHandle(&h[0], e->cur.p));
   CU(cudaIpcGetMemHandle(&h[1], e->pred.p));
   CU(cudaIpcGetMemHandle(&h[2], e->slab.flags.p));
   CU(cudaIpcGetMemHandle(&h[3], e->mm_cur.p));
   CU(cudaIpcGetMemHandle(&h[4], e->mm_pred.p));
   return 0;
}

int asd_slab_handle_bytes(void) { return (int)(5 * sizeof(cudaIpcMemHandle_t)); }

int asd_slab_connect_ipc(asd_engine* e, const void* lower, const void* upper) {
   Slab& sb = e->slab;
   if (!sb.on || !e->committed) return fail(-11, "slab: asd_set_slab + asd_commit first");
   CU(cudaSetDevice(e->device));
   const cudaIpcMemHandle_t* hs[2] = {(const cudaIpcMemHandle_t*)lower, (const cudaIpcMemHandle_t*)upper};
   // a reconnect (e.g. after a second asd_commit) closes the mappings of the previous connection first
   for (int q = 0; q < sb.n_opened; q++) cudaIpcCloseMemHandle(sb.opened[q]);
   sb.n_opened = 0;
   sb.connected = false;
   const bool same = memcmp(lower, upper, 5 * sizeof(cudaIpcMemHandle_t)) == 0;   // two slabs: one neighbour on both sides
   for (int side = 0; side < 2; side++) {
      if (side == 1 && same) {
         sb.peer_cur[1] = sb.peer_cur[0]; sb.peer_pred[1] = sb.peer_pred[0]; sb.peer_flags[1] = sb.peer_flags[0];
         sb.peer_mcur[1] = sb.peer_mcur[0]; sb.peer_mpred[1] = sb.peer_mpred[0];
         break;
      }
      void* p[5];
      for (int q = 0; q < 5; q++) {
         CU(cudaIpcOpenMemHandle(&p[q], hs[side][q], cudaIpcMemLazyEnablePeerAccess));
         if (sb.n_opened < 10) sb.opened[sb.n_opened++] = p[q];
      }
      sb.peer_cur[side] = (SpinVec*)p[0]; sb.peer_pred[side] = (SpinVec*)p[1]; sb.peer_flags[side] = (unsigned long long*)p[2];
      sb.peer_mcur[side] = (double*)p[3]; sb.peer_mpred[side] = (double*)p[4];
   }
   sb.connected = true;
   return 0;
}

int asd_slab_connect_local(asd_engine* e, asd_engine* lower, asd_engine* upper) {
   Slab& sb = e->slab;
   if (!sb.on || !e->committed) return fail(-11, "slab: asd_set_slab + asd_commit first");
   asd_engine* nb[2] = {lower, upper};
   for (int side = 0; side < 2; side++) {
      if (!nb[side]->slab.on || !nb[side]->committed || nb[side]->sd.Npad != e->sd.Npad) return fail(-11, "slab: neighbour engine is not a committed slab of the same shape");
      if (nb[side]->device != e->device) {
         int can = 0;
         CU(cudaDeviceCanAccessPeer(&can, e->device, nb[side]->device));
         if (!can) return fail(-11, "slab: no peer access between devices %d and %d", e->device, nb[side]->device);
         CU(cudaSetDevice(e->device));
         cudaError_t pe = cudaDeviceEnablePeerAccess(nb[side]->device, 0);
         if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CU(pe);
         cudaGetLastError();
      }
      sb.peer_cur[side] = nb[side]->cur.p; sb.peer_pred[side] = nb[side]->pred.p; sb.peer_flags[side] = nb[side]->slab.flags.p;
      sb.peer_mcur[side] = nb[side]->mm_cur.p; sb.peer_mpred[side] = nb[side]->mm_pred.p;
   }
   sb.connected = true;
   return 0;
}

int asd_slab_status(asd_engine* e, unsigned long long* epoch, int* error_flag) {
   CU(cudaSetDevice(e->device));
   if (epoch) *epoch = e->slab.epoch;
   if (error_flag) {
      *error_flag = 0;
      if (e->slab.on && e->slab.err.p) {
         CU(cudaStreamSynchronize(e->stream));
         CU(cudaMemcpy(error_flag, e->slab.err.p, sizeof(int), cudaMemcpyDeviceToHost));
         if (*error_flag) return fail(-12, "slab: timed out waiting for the halo of the %s neighbour", *error_flag == 1 ? "lower" : "upper");
      }
   }
   return 0;
}

int asd_layout_info(asd_engine* e, int* info5 /* 6 ints */) {
   if (!e->committed) return fail(-2, "asd_commit has not been called");
   const Tables& t = e->sd.t;
   info5[0] = t.staged; info5[1] = t.runs; info5[2] = t.ucap; info5[3] = t.runs ? t.union_max : 0;
   info5[4] = t.tile_slots;
   info5[5] = ((t.runs && (t.dm16 != nullptr || t.bq16 != nullptr)) ? 1 : 0) | ((t.runs && t.mm) ? 2 : 0);
   return 0;
}

long asd_launch_count(asd_engine* e) { return e->launches; }
int asd_synchronize(asd_engine* e) { CU(cudaSetDevice(e->device)); CU(cudaStreamSynchronize(e->stream)); return slab_check(e); }

}  // extern "C"

#include "asd_lattice_host.inl"
#include "asd_mc_block_host.inl"
#include "asd_legacy.inl"
