// Legacy native boundary (drop-in for source/gpu_files/fortranData.cpp + fort_helper.cpp + the time loop of
// cudaMdSimulation.cu:300-512).  Included by asd_engine.cu.
//
// Contract kept from the reference:
//  * fortrandata_set*_ store raw pointers into the Fortran module arrays; Fortran owns the storage;
//  * cudamdsim_initiateconstants_ dereferences the scalars and exits on an unsupported solver;
//  * cudamdsim_initiatematrices_ allocates device memory and uploads tables + state;
//  * cudamdsim_measurementphase_ runs nstep steps, calling back into the host's measurement routines with
//    mstep passed as size_t* (c_helper.h:30-36), and -- unlike the reference, which only copied state back as
//    a side effect of sampled measurements -- always writes the final emom/emomM/mmom/mmom2/mmomi/emom2 back
//    to the Fortran arrays before returning (uppasd.f90:344-350 writes the restart file from them).
// Deliberate differences: SDEalgh is honoured (1 = midpoint, 5 = Depondt; the reference CUDA path always ran
// Depondt -- set ASD_LEGACY_FORCE_DEPONDT=1 to mimic that).

#include <condition_variable>
#include <ctime>
#include <deque>
#include <mutex>
#include <thread>

extern "C" {
// gfortran-mangled host callbacks (c_helper.h:12-37); weak so that a non-Fortran host can register pointers instead
void __chelper_MOD_fortran_do_measurements(const size_t*, int*) __attribute__((weak));
void __chelper_MOD_fortran_measure_moment(const double*, const double*, const double*, const size_t*) __attribute__((weak));
void __chelper_MOD_fortran_flush_measurements(const size_t*) __attribute__((weak));
void __chelper_MOD_fortran_calc_simulation_status_variables(double*) __attribute__((weak));
}

namespace legacy {
struct FortranData {
   // constants (chelper.f90:171-173)
   char* stt = nullptr; int* SDEalgh = nullptr;
   unsigned int *rstep = nullptr, *nstep = nullptr, *Natom = nullptr, *Mensemble = nullptr, *max_no_neigh = nullptr;
   double *delta_t = nullptr, *gamma = nullptr, *k_bolt = nullptr, *mub = nullptr, *damping = nullptr, *binderc = nullptr, *mavg = nullptr;
   int* mompar = nullptr; char* initexc = nullptr;
   unsigned int *do_dm = nullptr, *max_no_dmneigh = nullptr, *do_jtensor = nullptr, *do_aniso = nullptr, *nHam = nullptr;
   // matrices (chelper.f90:181-184)
   double* ncoup = nullptr; unsigned int *nlist = nullptr, *nlistsize = nullptr;
   double *beff = nullptr, *b2eff = nullptr, *emomM = nullptr, *emom = nullptr, *emom2 = nullptr, *external_field = nullptr,
          *mmom = nullptr, *btorque = nullptr, *temperature = nullptr, *mmom0 = nullptr, *mmom2 = nullptr, *mmomi = nullptr, *dmvect = nullptr;
   unsigned int *dmlist = nullptr, *dmlistsize = nullptr;
   double *j_tensor = nullptr, *kaniso = nullptr, *eaniso = nullptr; unsigned int* taniso = nullptr; double* sb = nullptr; unsigned int* aHam = nullptr;
   // input data (chelper.f90:186)
   int *gpu_mode = nullptr, *gpu_rng = nullptr, *gpu_rng_seed = nullptr;
   // extras (new)
   double *Landeg = nullptr, *lambda1_array = nullptr, *temprescale = nullptr;
   unsigned int *do_bq = nullptr, *nn_bq_tot = nullptr, *bqlist = nullptr, *bqlistsize = nullptr; double* j_bq = nullptr;
   // supercell shape (new, optional)
   unsigned int *NA = nullptr, *N1 = nullptr, *N2 = nullptr, *N3 = nullptr; char *BC1 = nullptr, *BC2 = nullptr, *BC3 = nullptr;
};
static FortranData fd;
static asd_engine* eng = nullptr;
static bool constants_ok = false, matrices_ok = false;
static asd_cb_do_measurements cb_do = nullptr;
static asd_cb_measure_moment cb_measure = nullptr;
static asd_cb_flush_measurements cb_flush = nullptr;
static asd_cb_status cb_status = nullptr;

static int do_measurements(size_t mstep) {
   int do_copy = 0;
   if (cb_do) cb_do(&mstep, &do_copy);
   else if (__chelper_MOD_fortran_do_measurements) __chelper_MOD_fortran_do_measurements(&mstep, &do_copy);
   return do_copy;
}
static void measure_moment(size_t mstep) {
   if (cb_measure) cb_measure(fd.emomM, fd.emom, fd.mmom, &mstep);
   else if (__chelper_MOD_fortran_measure_moment) __chelper_MOD_fortran_measure_moment(fd.emomM, fd.emom, fd.mmom, &mstep);
}
static void flush_measurements(size_t mstep) {
   if (cb_flush) cb_flush(&mstep);
   else if (__chelper_MOD_fortran_flush_measurements) __chelper_MOD_fortran_flush_measurements(&mstep);
}
static void status(double* mavg) {
   if (cb_status) cb_status(mavg);
   else if (__chelper_MOD_fortran_calc_simulation_status_variables) __chelper_MOD_fortran_calc_simulation_status_variables(mavg);
}
[[noreturn]] static void die(const char* what) {
   std::fprintf(stderr, "uppasd_b200: %s: %s\n", what, asd_last_error());
   std::exit(EXIT_FAILURE);
}
static void copy_to_fortran(bool all) {
   if (asd_get_moments(eng, fd.emom, fd.emomM, fd.mmom)) die("copy to host");
   if (all) {
      const size_t NM = (size_t)eng->N * eng->M;
      for (size_t q = 0; q < NM; q++) {
         if (fd.mmom2) fd.mmom2[q] = fd.mmom[q];
         if (fd.mmomi) fd.mmomi[q] = 1.0 / fd.mmom[q];
      }
      if (fd.emom2) std::memcpy(fd.emom2, fd.emom, 3 * NM * sizeof(double));
   }
}

// ---- asynchronous measurement path (replaces gpu_files/cudaMeasurement.cu:109-182 + measurementQueue.cpp:63-121) ----
// A sampled step must hand emomM / emom / mmom of THAT step to the host's measurement routine.  Instead of stopping the time
// loop for the 56 N M bytes, the step's state is unpacked into one of NSLOT device staging sets on the compute stream, an event
// hands it to a copy stream that lands it in pinned host memory, and a worker thread calls fortran_measure_moment on the
// landed buffers while the compute stream is already integrating the following steps.  The host routine receives the buffers
// as its arguments (chelper.f90:106-126 measures ext_emomM / ext_emom / ext_mmom), so the Fortran module arrays are not touched
// until the final copy-back.  ASD_LEGACY_SYNC=1 keeps the blocking copy into the module arrays (needed by a host whose
// measurement reads the module arrays themselves, e.g. do_sc through correlation_wrapper).
struct SnapshotRing {
   static constexpr int NSLOT = 3;
   struct Slot { double *d_e = nullptr, *d_eM = nullptr, *d_m = nullptr, *h_e = nullptr, *h_eM = nullptr, *h_m = nullptr; cudaEvent_t packed = nullptr, landed = nullptr; };
   Slot slot[NSLOT];
   cudaStream_t copy = nullptr;
   size_t NM = 0;
   int next = 0;
   bool ready = false, stop = false;
   std::thread worker;
   std::mutex mu;
   std::condition_variable cv, cv_free;
   std::deque<std::pair<int, size_t>> queue;      // (slot, mstep) in sampling order
   int busy[NSLOT] = {0, 0, 0};
   long served = 0;

   int init(asd_engine* e) {
      release();
      NM = (size_t)e->N * e->M;
      if (cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking) != cudaSuccess) return -1;
      for (auto& s : slot) {
         if (cudaMalloc((void**)&s.d_e, 3 * NM * sizeof(double)) != cudaSuccess || cudaMalloc((void**)&s.d_eM, 3 * NM * sizeof(double)) != cudaSuccess ||
             cudaMalloc((void**)&s.d_m, NM * sizeof(double)) != cudaSuccess || cudaMallocHost((void**)&s.h_e, 3 * NM * sizeof(double)) != cudaSuccess ||
             cudaMallocHost((void**)&s.h_eM, 3 * NM * sizeof(double)) != cudaSuccess || cudaMallocHost((void**)&s.h_m, NM * sizeof(double)) != cudaSuccess ||
             cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming) != cudaSuccess) {
            release();
            return -1;
         }
      }
      stop = false;
      worker = std::thread([this]() { this->run(); });
      ready = true;
      return 0;
   }
   void run() {
      for (;;) {
         std::pair<int, size_t> job;
         {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this]() { return stop || !queue.empty(); });
            if (queue.empty()) return;
            job = queue.front();
            queue.pop_front();
         }
         Slot& s = slot[job.first];
         cudaEventSynchronize(s.landed);
         const size_t mstep = job.second;
         if (cb_measure) cb_measure(s.h_eM, s.h_e, s.h_m, &mstep);
         else if (__chelper_MOD_fortran_measure_moment) __chelper_MOD_fortran_measure_moment(s.h_eM, s.h_e, s.h_m, &mstep);
         {
            std::lock_guard<std::mutex> lk(mu);
            busy[job.first] = 0;
            served++;
         }
         cv_free.notify_all();
      }
   }
   // enqueue the state of the engine's stream position as the sample of step mstep
   int sample(asd_engine* e, size_t mstep) {
      const int q = next;
      next = (next + 1) % NSLOT;
      {
         std::unique_lock<std::mutex> lk(mu);
         cv_free.wait(lk, [&]() { return busy[q] == 0; });       // the worker is done with this slot's previous sample
         busy[q] = 1;
      }
      Slot& s = slot[q];
      if (ensure_layout(e, e->state_layout == 2 ? 2 : 1)) return -1;      // the first sample precedes the first step: state not uploaded yet
      Layout& L = (e->state_layout == 2) ? e->mc : e->sd;
      dim3 g, b;
      launch_cfg(L.Npad, e->M, g, b);
      unpack_kernel<<<g, b, 0, e->stream>>>(e->N, L.Npad, e->M, L.d_orig.p, e->cur.p, s.d_e, s.d_eM, s.d_m);
      e->launches++;
      if (cudaGetLastError() != cudaSuccess) return -1;
      cudaEventRecord(s.packed, e->stream);
      cudaStreamWaitEvent(copy, s.packed, 0);
      cudaMemcpyAsync(s.h_e, s.d_e, 3 * NM * sizeof(double), cudaMemcpyDeviceToHost, copy);
      cudaMemcpyAsync(s.h_eM, s.d_eM, 3 * NM * sizeof(double), cudaMemcpyDeviceToHost, copy);
      cudaMemcpyAsync(s.h_m, s.d_m, NM * sizeof(double), cudaMemcpyDeviceToHost, copy);
      cudaEventRecord(s.landed, copy);
      {
         std::lock_guard<std::mutex> lk(mu);
         queue.emplace_back(q, mstep);
      }
      cv.notify_one();
      return 0;
   }
   void drain() {
      if (!ready) return;
      std::unique_lock<std::mutex> lk(mu);
      cv_free.wait(lk, [this]() { return queue.empty() && busy[0] == 0 && busy[1] == 0 && busy[2] == 0; });
   }
   ~SnapshotRing() { if (worker.joinable()) { { std::lock_guard<std::mutex> lk(mu); stop = true; queue.clear(); } cv.notify_all(); worker.detach(); } }
   void release() {
      if (worker.joinable()) {
         { std::lock_guard<std::mutex> lk(mu); stop = true; }
         cv.notify_all();
         worker.join();
      }
      for (auto& s : slot) {
         if (s.d_e) cudaFree(s.d_e); if (s.d_eM) cudaFree(s.d_eM); if (s.d_m) cudaFree(s.d_m);
         if (s.h_e) cudaFreeHost(s.h_e); if (s.h_eM) cudaFreeHost(s.h_eM); if (s.h_m) cudaFreeHost(s.h_m);
         if (s.packed) cudaEventDestroy(s.packed); if (s.landed) cudaEventDestroy(s.landed);
         s = Slot();
      }
      if (copy) { cudaStreamDestroy(copy); copy = nullptr; }
      queue.clear();
      busy[0] = busy[1] = busy[2] = 0;
      ready = false;
   }
};
static SnapshotRing ring;
}  // namespace legacy

extern "C" {

/* samples served by the asynchronous measurement path so far (0: the blocking path is in use) */
long asd_legacy_async_samples(void) { return legacy::ring.served; }

void fortrandata_setconstants_(char* p1, int* p2, unsigned int* p3, unsigned int* p4, unsigned int* p5, unsigned int* p6,
                               unsigned int* p7, double* p8, double* p9, double* p10, double* p11, double* p12,
                               double* p13, double* p14, int* p15, char* p16, unsigned int* p17, unsigned int* p18,
                               unsigned int* p19, unsigned int* p20, unsigned int* p21) {
   auto& f = legacy::fd;
   f.stt = p1; f.SDEalgh = p2; f.rstep = p3; f.nstep = p4; f.Natom = p5; f.Mensemble = p6; f.max_no_neigh = p7;
   f.delta_t = p8; f.gamma = p9; f.k_bolt = p10; f.mub = p11; f.damping = p12; f.binderc = p13; f.mavg = p14;
   f.mompar = p15; f.initexc = p16; f.do_dm = p17; f.max_no_dmneigh = p18; f.do_jtensor = p19; f.do_aniso = p20; f.nHam = p21;
}

void fortrandata_setmatrices_(double* p1, unsigned int* p2, unsigned int* p3, double* p4, double* p5, double* p6, double* p7,
                              double* p8, double* p9, double* p10, double* p11, double* p12, double* p13, double* p14,
                              double* p15, double* p16, unsigned int* p17, unsigned int* p18, double* p19, double* p20,
                              double* p21, unsigned int* p22, double* p23, unsigned int* p24) {
   auto& f = legacy::fd;
   f.ncoup = p1; f.nlist = p2; f.nlistsize = p3; f.beff = p4; f.b2eff = p5; f.emomM = p6; f.emom = p7; f.emom2 = p8;
   f.external_field = p9; f.mmom = p10; f.btorque = p11; f.temperature = p12; f.mmom0 = p13; f.mmom2 = p14; f.mmomi = p15;
   f.dmvect = p16; f.dmlist = p17; f.dmlistsize = p18; f.j_tensor = p19; f.kaniso = p20; f.eaniso = p21; f.taniso = p22;
   f.sb = p23; f.aHam = p24;
}

void fortrandata_setinputdata_(int* p1, int* p2, int* p3) {
   legacy::fd.gpu_mode = p1; legacy::fd.gpu_rng = p2; legacy::fd.gpu_rng_seed = p3;
}

void fortrandata_setextras_(double* Landeg, double* lambda1_array, double* temprescale, unsigned int* do_bq,
                            unsigned int* nn_bq_tot, unsigned int* bqlist, unsigned int* bqlistsize, double* j_bq) {
   auto& f = legacy::fd;
   f.Landeg = Landeg; f.lambda1_array = lambda1_array; f.temprescale = temprescale;
   f.do_bq = do_bq; f.nn_bq_tot = nn_bq_tot; f.bqlist = bqlist; f.bqlistsize = bqlistsize; f.j_bq = j_bq;
}

void fortrandata_setlattice_(unsigned int* NA, unsigned int* N1, unsigned int* N2, unsigned int* N3, char* BC1, char* BC2, char* BC3) {
   auto& f = legacy::fd;
   f.NA = NA; f.N1 = N1; f.N2 = N2; f.N3 = N3; f.BC1 = BC1; f.BC2 = BC2; f.BC3 = BC3;
}

asd_engine* asd_legacy_engine(void) { return legacy::eng; }

void asd_set_callbacks(asd_cb_do_measurements a, asd_cb_measure_moment b, asd_cb_flush_measurements c, asd_cb_status d) {
   legacy::cb_do = a; legacy::cb_measure = b; legacy::cb_flush = c; legacy::cb_status = d;
}

void cudamdsim_initiateconstants_(void) {
   using namespace legacy;
   if (!fd.SDEalgh || !fd.Natom) { std::fprintf(stderr, "uppasd_b200: fortrandata_setconstants_ has not been called\n"); std::exit(EXIT_FAILURE); }
   const int alg = *fd.SDEalgh;
   // the reference accepts 1, 4, 5, 11 and always integrates with Depondt (cudaMdSimulation.cu:35-38,319);
   // this build honours the two solvers on its path and refuses the rest the same way the reference does
   if (!(alg == 1 || alg == 5)) { std::fprintf(stderr, "Invalid SDEalgh!\n"); std::exit(EXIT_FAILURE); }
   if (fd.gpu_rng && (*fd.gpu_rng < 0 || *fd.gpu_rng > 3)) { std::fprintf(stderr, "Unknown gpu_rng %d\n", *fd.gpu_rng); std::exit(EXIT_FAILURE); }
   if (eng) { asd_destroy(eng); eng = nullptr; }
   if (asd_create(&eng, -1)) die("cudamdsim_initiateconstants_");
   asd_set_constants(eng, *fd.gamma, *fd.k_bolt, *fd.mub, eng->mry);
   constants_ok = true; matrices_ok = false;
}

void cudamdsim_initiatematrices_(void) {
   using namespace legacy;
   if (!constants_ok) { std::fprintf(stderr, "uppasd_b200: constants not initiated!\n"); std::exit(EXIT_FAILURE); }
   const int N = (int)*fd.Natom, M = (int)*fd.Mensemble, NH = (int)*fd.nHam;
   if (asd_set_system(eng, N, M, NH, (const int*)fd.aHam)) die("set_system");
   if (fd.NA && fd.N1 && fd.N2 && fd.N3 && fd.BC1 && fd.BC2 && fd.BC3) {
      const char bc[4] = {*fd.BC1, *fd.BC2, *fd.BC3, 0};
      if (asd_set_lattice_hint(eng, (int)*fd.NA, (int)*fd.N1, (int)*fd.N2, (int)*fd.N3, bc)) die("set_lattice_hint");
   }
   if (fd.do_jtensor && *fd.do_jtensor == 1) {
      // tensorial exchange (cudaHamiltonianCalculations.cu:274-343 in the reference's native path): j_tens(3,3,z,nHam)
      if (asd_set_jtensor(eng, (int)*fd.max_no_neigh, (const int*)fd.nlist, (const int*)fd.nlistsize, fd.j_tensor)) die("set_jtensor");
   } else if (asd_set_exchange(eng, (int)*fd.max_no_neigh, (const int*)fd.nlist, (const int*)fd.nlistsize, fd.ncoup)) die("set_exchange");
   if (fd.do_dm && *fd.do_dm == 1)
      if (asd_set_dm(eng, (int)*fd.max_no_dmneigh, (const int*)fd.dmlist, (const int*)fd.dmlistsize, fd.dmvect)) die("set_dm");
   if (fd.do_bq && *fd.do_bq == 1)
      if (asd_set_bq(eng, (int)*fd.nn_bq_tot, (const int*)fd.bqlist, (const int*)fd.bqlistsize, fd.j_bq)) die("set_bq");
   if (fd.do_aniso && *fd.do_aniso != 0)
      if (asd_set_anisotropy(eng, (const int*)fd.taniso, fd.eaniso, fd.kaniso, fd.sb)) die("set_anisotropy");
   if (asd_set_external_field(eng, fd.external_field)) die("set_external_field");
   if (fd.stt && *fd.stt != 'N' && fd.btorque) if (asd_set_torque(eng, fd.btorque)) die("set_torque");
   std::vector<double> lam(N, *fd.damping);
   int alg = *fd.SDEalgh;
   const char* force = std::getenv("ASD_LEGACY_FORCE_DEPONDT");
   if (force && force[0] == '1') alg = 5;
   unsigned long long seed = (fd.gpu_rng_seed && *fd.gpu_rng_seed != 0) ? (unsigned long long)*fd.gpu_rng_seed : (unsigned long long)std::time(nullptr);
   if (asd_set_llg(eng, alg, *fd.delta_t, fd.Landeg, fd.lambda1_array ? fd.lambda1_array : lam.data(), fd.temperature,
                   fd.temprescale ? *fd.temprescale : 1.0, *fd.mompar, seed)) die("set_llg");
   if (asd_set_moments(eng, fd.emom, fd.mmom, fd.mmom0)) die("set_moments");
   if (asd_commit(eng)) {
      // like the reference (cudaMdSimulation.cu:171-181): report and leave the state un-initiated
      std::fprintf(stderr, "uppasd_b200: initiateMatrices failed: %s\n", asd_last_error());
      matrices_ok = false;
      return;
   }
   matrices_ok = true;
}

void cudamdsim_measurementphase_(void) {
   using namespace legacy;
   std::setbuf(stdout, nullptr);
   std::printf("uppasd_b200: md simulation starting\n");
   if (!matrices_ok) { std::fprintf(stderr, "uppasd_b200: not initiated!\n"); return; }
   const size_t rstep = *fd.rstep, nstep = *fd.nstep;
   const char* senv = std::getenv("ASD_LEGACY_SYNC");
   bool async = !(senv && senv[0] == '1');
   if (async && ring.init(eng) != 0) async = false;       // no room for the staging ring: blocking copies
   size_t pending_first = 0, pending = 0;  // steps enqueued lazily so that runs of unsampled steps cost no host sync
   auto flush_steps = [&]() {
      if (pending) { if (asd_sd_steps(eng, (long)pending, (long)pending_first)) die("sd_steps"); pending = 0; }
   };
   std::vector<double> msum((size_t)3 * eng->M);
   for (size_t mstep = rstep + 1; mstep <= rstep + nstep; mstep++) {
      if (do_measurements(mstep)) {
         flush_steps();
         if (async) { if (ring.sample(eng, mstep)) die("measurement snapshot"); }
         else { copy_to_fortran(false); measure_moment(mstep); }
      }
      // progress line every 5 % (cudaMdSimulation.cu:283-297).  Mbar = calc_mavrg of the current state (prn_averages.f90:414-456):
      // mean over the ensembles of |sum_i emomM| / Natom, from the per-tile sums the corrector launch leaves on the device
      // (24 M bytes to the host instead of the 56 N M bytes the reference copies for fortran_calc_simulation_status_variables).
      if (nstep > 20 ? (mstep % ((rstep + nstep) / 20) == 0) : true) {
         flush_steps();
         if (async) {
            if (asd_measure(eng, msum.data(), nullptr)) die("status");
            double mb = 0.0;
            for (int k = 0; k < eng->M; k++)
               mb += std::sqrt(msum[3 * k] * msum[3 * k] + msum[3 * k + 1] * msum[3 * k + 1] + msum[3 * k + 2] * msum[3 * k + 2]) / eng->N;
            if (fd.mavg) *fd.mavg = mb / eng->M;
         } else {
            copy_to_fortran(false);
            if (fd.mavg) status(fd.mavg);
         }
         if (nstep > 20) std::printf("CUDA: %3ld%% done. Mbar: %10.6f. U: %8.5f.\n", (long)(mstep * 100 / (rstep + nstep)), fd.mavg ? *fd.mavg : 0.0, fd.binderc ? *fd.binderc : 0.0);
         else std::printf("CUDA: Iteration %ld Mbar %13.6f\n", (long)mstep, fd.mavg ? *fd.mavg : 0.0);
      }
      if (!pending) pending_first = mstep;
      pending++;
   }
   flush_steps();
   if (async) ring.drain();                 // every sampled step has been measured, in order, before the final one
   const size_t last = rstep + nstep + 1;
   copy_to_fortran(true);
   if (do_measurements(last)) measure_moment(last);
   flush_measurements(last);
   if (asd_synchronize(eng)) die("synchronize");
   if (async) ring.release();
}

// ---- new sibling entries in the same F77 style (boundaries the reference never had) ----------------
// SD initial phase: one phase of sd_iphase (sd_driver.f90:144-289) on the engine initiated by
// cudamdsim_initiatematrices_: nstep steps at temperature Temp with time step delta_t and damping lambda1,
// solver ipSDEalgh (1 or 5).  The state stays on the device; emom/emomM/mmom are written back on return.
void cudamdsim_initialphase_(unsigned int* ipnstep, double* ipTemp, double* ipdelta_t, double* iplambda1, int* ipSDEalgh,
                             unsigned int* first_step) {
   using namespace legacy;
   if (!matrices_ok) { std::fprintf(stderr, "uppasd_b200: not initiated!\n"); return; }
   const int N = eng->N;
   std::vector<double> lam(N, *iplambda1), temp(N, *ipTemp);
   if (asd_set_llg(eng, *ipSDEalgh, *ipdelta_t, fd.Landeg, lam.data(), temp.data(), fd.temprescale ? *fd.temprescale : 1.0,
                   *fd.mompar, eng->seed)) die("cudamdsim_initialphase_");
   if (asd_sd_steps(eng, (long)*ipnstep, (long)(first_step ? *first_step : 1))) die("cudamdsim_initialphase_");
   copy_to_fortran(true);
   // restore the measurement-phase parameters
   std::vector<double> lam0(N, *fd.damping);
   if (asd_set_llg(eng, eng->SDEalgh, *fd.delta_t, fd.Landeg, fd.lambda1_array ? fd.lambda1_array : lam0.data(), fd.temperature,
                   fd.temprescale ? *fd.temprescale : 1.0, *fd.mompar, eng->seed)) die("cudamdsim_initialphase_");
}

// Monte Carlo: the body of the sweep loops of mc_iphase / mc_mphase / mc_minimal (mc_driver.f90:117-201, 310-430,
// 516-525): nsweeps calls of mc_evolve(..., Temp, temprescale, mode, ..., emomM, emom, mmom, ..., extfield, ...) on the
// tables handed over by fortrandata_setmatrices_.  emom / mmom are read from the host arrays when *upload != 0 (the
// host changed them since the last call) and emom / emomM / mmom are written back on return (mc_driver.f90:215-216).
void cudamcsim_evolve_(char* mode, unsigned int* nsweeps, unsigned int* first_sweep, double* Temp, double* temprescale,
                       double* extfield, int* upload) {
   using namespace legacy;
   if (!matrices_ok) { std::fprintf(stderr, "uppasd_b200: not initiated!\n"); return; }
   if (upload && *upload) { if (asd_set_moments(eng, fd.emom, fd.mmom, fd.mmom0)) die("cudamcsim_evolve_"); }
   if (asd_mc_sweeps(eng, *mode, (long)*nsweeps, (long)*first_sweep, *Temp, temprescale ? *temprescale : 1.0, extfield)) die("cudamcsim_evolve_");
   copy_to_fortran(true);
}

// ---- pyasd's second caller (source/pyasd.f90): the bind(c) entry points the Python driver (uppasd.pyasd) binds, acting on
//      the engine and the module arrays that FortranData_Initiate handed over.  Same names, argument order and meaning; every
//      argument by reference, characters as char*.
// relax_  (pyasd.f90:255-298): imode 'M' / 'H' -> mc_minimal: instep sweeps of mc_evolve at itemperature (mc_driver.f90:457-541);
//         otherwise damping = idamping and sd_minimal(emomM, emom, mmom, instep, 1, itemperature): instep midpoint steps at
//         itemperature with the module time step (sd_driver.f90:1162-1340).  itimestep is accepted and NOT used, exactly like the
//         reference.  moments(3, natom, mensemble) = emomM on return; the module arrays emom / emomM / mmom are updated too.
static long pyasd_step = 1;      // noise / draw counter: consecutive relax_ calls continue one random stream
void relax_(double* moments, int* natom, int* mensemble, char* imode, int* instep, double* itemperature, double* itimestep,
            double* idamping) {
   using namespace legacy;
   (void)itimestep;
   if (!matrices_ok) { std::fprintf(stderr, "uppasd_b200: relax_: not initiated!\n"); return; }
   if (*natom != eng->N || *mensemble != eng->M) { std::fprintf(stderr, "uppasd_b200: relax_: shape (%d, %d) is not the engine's (%d, %d)\n", *natom, *mensemble, eng->N, eng->M); std::exit(EXIT_FAILURE); }
   const int N = eng->N, n = std::max(*instep, 0);
   if (*imode == 'M' || *imode == 'H') {
      // mc_evolve's extfield argument is hfield(1:3); the handed-over external_field array holds it for every atom
      const double ext[3] = {fd.external_field ? fd.external_field[0] : 0.0, fd.external_field ? fd.external_field[1] : 0.0,
                             fd.external_field ? fd.external_field[2] : 0.0};
      if (asd_mc_sweeps(eng, *imode, n, pyasd_step, *itemperature, fd.temprescale ? *fd.temprescale : 1.0, ext)) die("relax_");
   } else {
      std::vector<double> lam(N, *idamping), temp(N, *itemperature);
      if (asd_set_llg(eng, 1, *fd.delta_t, fd.Landeg, lam.data(), temp.data(), fd.temprescale ? *fd.temprescale : 1.0, *fd.mompar, eng->seed)) die("relax_");
      if (asd_sd_steps(eng, n, pyasd_step)) die("relax_");
      // back to the measurement-phase parameters (damping1 stays idamping in the reference; the legacy loop re-reads *fd.damping)
      std::vector<double> lam0(N, *fd.damping);
      if (asd_set_llg(eng, eng->SDEalgh, *fd.delta_t, fd.Landeg, fd.lambda1_array ? fd.lambda1_array : lam0.data(), fd.temperature,
                      fd.temprescale ? *fd.temprescale : 1.0, *fd.mompar, eng->seed)) die("relax_");
   }
   pyasd_step += n;
   copy_to_fortran(true);
   std::memcpy(moments, fd.emomM, (size_t)3 * N * eng->M * sizeof(double));
}

// get_emom_ (pyasd.f90:316-328): moments = emom
void get_emom_(double* moments, int* natom, int* mensemble) {
   using namespace legacy;
   if (!matrices_ok || *natom != eng->N || *mensemble != eng->M) { std::fprintf(stderr, "uppasd_b200: get_emom_: not initiated or wrong shape\n"); return; }
   if (asd_get_moments(eng, moments, nullptr, nullptr)) die("get_emom_");
}

// put_emom_ (pyasd.f90:330-350): emom = moments, emom2 = emom, emomM = moments * mmom
void put_emom_(const double* moments, int* natom, int* mensemble) {
   using namespace legacy;
   if (!matrices_ok || *natom != eng->N || *mensemble != eng->M) { std::fprintf(stderr, "uppasd_b200: put_emom_: not initiated or wrong shape\n"); return; }
   const size_t NM = (size_t)eng->N * eng->M;
   if (asd_get_moments(eng, nullptr, nullptr, fd.mmom)) die("put_emom_");      // the magnitudes the device holds (mompar may have changed them)
   if (fd.emom != moments) std::memcpy(fd.emom, moments, 3 * NM * sizeof(double));
   if (fd.emom2) std::memcpy(fd.emom2, moments, 3 * NM * sizeof(double));
   for (size_t q = 0; q < NM; q++)
      for (int a = 0; a < 3; a++) fd.emomM[3 * q + a] = moments[3 * q + a] * fd.mmom[q];
   if (asd_set_moments(eng, moments, fd.mmom, fd.mmom0)) die("put_emom_");
}

// get_beff_ (pyasd.f90:356-369): call effective_field(); fields = beff
void get_beff_(double* fields, int* natom, int* mensemble) {
   using namespace legacy;
   if (!matrices_ok || *natom != eng->N || *mensemble != eng->M) { std::fprintf(stderr, "uppasd_b200: get_beff_: not initiated or wrong shape\n"); return; }
   if (asd_effective_field(eng, fields, nullptr, nullptr, nullptr)) die("get_beff_");
   if (fd.beff && fd.beff != fields) std::memcpy(fd.beff, fields, (size_t)3 * eng->N * eng->M * sizeof(double));
}

// get_energy_ (pyasd.f90:505-517): call effective_field(energy); energy = energy / (Natom * Mensemble)
void get_energy_(double* energy) {
   using namespace legacy;
   if (!matrices_ok) { std::fprintf(stderr, "uppasd_b200: get_energy_: not initiated!\n"); return; }
   std::vector<double> en(eng->M, 0.0);
   if (asd_effective_field(eng, nullptr, nullptr, nullptr, en.data())) die("get_energy_");
   double tot = 0.0;
   for (double v : en) tot += v;
   *energy = tot / ((double)eng->N * eng->M);
}

void cmdsim_initiateconstants_(void) { cudamdsim_initiateconstants_(); }
void cmdsim_initiatefortran_(void) { cudamdsim_initiatematrices_(); }
void cmdsim_measurementphase_(void) { cudamdsim_measurementphase_(); }

}  // extern "C"
