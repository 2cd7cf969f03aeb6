// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-time-step hot path.
// Nothing under uppasd_b200/ may include, link or call this file.
//
// Follows the Fortran loop by loop (paths relative to the reference root):
//   effective field   source/Hamiltonian/hamiltonianactions.f90:108-252 (effective_field_full),
//                     :431-466 heisenberg_field, :546-580 dzyaloshinskii_moriya_field,
//                     :775-796 biquadratic_field, :842-879 uniaxial, :883-919 cubic
//   midpoint solver   source/Evolution/midpoint.f90:34-181 (smodeulermpt), :188-323 (modeulermpf)
//   Depondt solver    source/Evolution/depondt.f90:50-193, :203-333
//   moment update     source/Evolution/updatemoments.f90:19-145 (calcm, copym)
//   noise amplitude   source/RNG/randomnumbers.f90:625-771 (rannum, llg=1)
//   Monte Carlo       source/MonteCarlo/montecarlo.f90:44-273 (mc_evolve), 277-301 (visiting order),
//                     source/MonteCarlo/montecarlo_common.f90:25-79 (trial move), 431-865 (dE),
//                     139-200 (Metropolis accept), 371-422 (heat bath)
//   observables       source/Measurement/prn_averages.f90:414-456 (sum of moments), 919-1034 (cumulants)
//
// Arrays keep Fortran shapes (column-major): emom(3,N,M), mmom(N,M), nlist(z,N) 1-based, ncoup(z,NH).
// The OpenMP schedule mirrors the reference: collapse over (ensemble, atom), static.
// Compile with -ffp-contract=off so every product/sum rounds as written in the Fortran.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

// thread count of the OpenMP loops (a launcher such as torchrun exports OMP_NUM_THREADS=1 to its workers; the timed CPU
// baseline must use the host cores it reports): returns the count now in force
int orc_set_num_threads(int n) {
#ifdef _OPENMP
   if (n > 0) omp_set_num_threads(n);
   return omp_get_max_threads();
#else
   (void)n;
   return 1;
#endif
}

struct OrcHam {
   int Natom, Mensemble, nHam;
   // exchange
   int max_no_neigh;
   const int* nlist;      // (max_no_neigh, Natom)
   const int* nlistsize;  // (nHam)
   const double* ncoup;   // (max_no_neigh, nHam)
   const int* aHam;       // (Natom)
   // DM
   int do_dm, max_no_dmneigh;
   const int* dmlist;      // (max_no_dmneigh, Natom)
   const int* dmlistsize;  // (nHam)
   const double* dm_vect;  // (3, max_no_dmneigh, nHam)
   // BQ
   int do_bq, nn_bq_tot;
   const int* bqlist;      // (nn_bq_tot, Natom)
   const int* bqlistsize;  // (nHam)
   const double* j_bq;     // (nn_bq_tot, nHam)
   // anisotropy
   int do_anisotropy;
   const int* taniso;      // (Natom)
   const double* eaniso;   // (3, Natom)
   const double* kaniso;   // (2, Natom)
   const double* sb;       // (Natom)
   // tensorial exchange: ncoup holds j_tens(3,3,max_no_neigh,nHam)
   int do_jtensor;
};

// ---- field terms -----------------------------------------------------------------------------
static inline void heisenberg_field(const OrcHam& H, long i, long k, const double* emomM, double* f) {
   const int ih = H.aHam[i - 1];
   const int z = H.max_no_neigh;
   const double* eM = emomM + 3 * (size_t)H.Natom * (k - 1);
   for (int j = 1; j <= H.nlistsize[ih - 1]; j++) {
      const double c = H.ncoup[(j - 1) + (size_t)z * (ih - 1)];
      const long nb = H.nlist[(j - 1) + (size_t)z * (i - 1)];
      f[0] = f[0] + c * eM[3 * (nb - 1)];
      f[1] = f[1] + c * eM[3 * (nb - 1) + 1];
      f[2] = f[2] + c * eM[3 * (nb - 1) + 2];
   }
}

// tensor_field (hamiltonianactions.f90:499-542): field += J(:,1) m_x + J(:,2) m_y + J(:,3) m_z.  The Fortran indexes
// j_tens(:,:,j,i) with the ATOM number; this restatement uses the Hamiltonian row aHam(i), identical for do_reduced N
// (the only mode the reference's own tensor tests use) and in-bounds for do_reduced Y.
static inline void tensor_field(const OrcHam& H, long i, long k, const double* emomM, double* f) {
   const int ih = H.aHam[i - 1];
   const int z = H.max_no_neigh;
   const double* eM = emomM + 3 * (size_t)H.Natom * (k - 1);
   for (int j = 1; j <= H.nlistsize[ih - 1]; j++) {
      const double* J = H.ncoup + 9 * ((j - 1) + (size_t)z * (ih - 1));   // J(a,b) at a + 3 b
      const long nb = H.nlist[(j - 1) + (size_t)z * (i - 1)];
      const double mx = eM[3 * (nb - 1)], my = eM[3 * (nb - 1) + 1], mz = eM[3 * (nb - 1) + 2];
      for (int a = 0; a < 3; a++) f[a] = f[a] + J[a] * mx + J[a + 3] * my + J[a + 6] * mz;
   }
}

static inline void dm_field(const OrcHam& H, long i, long k, const double* emomM, double* f) {
   const int ih = H.aHam[i - 1];
   const int z = H.max_no_dmneigh;
   const double* eM = emomM + 3 * (size_t)H.Natom * (k - 1);
   for (int j = 1; j <= H.dmlistsize[ih - 1]; j++) {
      const double* D = H.dm_vect + 3 * ((j - 1) + (size_t)z * (ih - 1));
      const long nb = H.dmlist[(j - 1) + (size_t)z * (i - 1)];
      const double mx = eM[3 * (nb - 1)], my = eM[3 * (nb - 1) + 1], mz = eM[3 * (nb - 1) + 2];
      f[0] = f[0] + D[2] * my - D[1] * mz;
      f[1] = f[1] + D[0] * mz - D[2] * mx;
      f[2] = f[2] + D[1] * mx - D[0] * my;
   }
}

static inline void bq_field(const OrcHam& H, long i, long k, const double* emomM, double* f) {
   const int ih = H.aHam[i - 1];
   const int z = H.nn_bq_tot;
   const double* eM = emomM + 3 * (size_t)H.Natom * (k - 1);
   const double* mi = eM + 3 * (i - 1);
   for (int j = 1; j <= H.bqlistsize[ih - 1]; j++) {
      const long nb = H.bqlist[(j - 1) + (size_t)z * (i - 1)];
      const double* mj = eM + 3 * (nb - 1);
      const double dot = mj[0] * mi[0] + mj[1] * mi[1] + mj[2] * mi[2];
      const double c = 2.0 * H.j_bq[(j - 1) + (size_t)z * (ih - 1)] * dot;
      f[0] = f[0] + c * mj[0];
      f[1] = f[1] + c * mj[1];
      f[2] = f[2] + c * mj[2];
   }
}

static inline void uniaxial_field(const OrcHam& H, long i, const double* m, double* f) {
   const double* e = H.eaniso + 3 * (i - 1);
   const double* kk = H.kaniso + 2 * (i - 1);
   const double tt1 = m[0] * e[0] + m[1] * e[1] + m[2] * e[2];
   const double tt2 = kk[0] + 2.0 * kk[1] * (1.0 - tt1 * tt1);
   const double tt3 = 2.0 * tt1 * tt2;
   f[0] = f[0] - tt3 * e[0];
   f[1] = f[1] - tt3 * e[1];
   f[2] = f[2] - tt3 * e[2];
}

static inline void cubic_field(const OrcHam& H, long i, const double* m, double* f) {
   const double* kk = H.kaniso + 2 * (i - 1);
   const double x2 = m[0] * m[0], y2 = m[1] * m[1], z2 = m[2] * m[2];
   f[0] = f[0] + 2.0 * kk[0] * m[0] * (y2 + z2) + 2.0 * kk[1] * m[0] * (y2 * z2);
   f[1] = f[1] + 2.0 * kk[0] * m[1] * (z2 + x2) + 2.0 * kk[1] * m[1] * (z2 * x2);
   f[2] = f[2] + 2.0 * kk[0] * m[2] * (x2 + y2) + 2.0 * kk[1] * m[2] * (x2 * y2);
}

// one site (hamiltonianactions.f90:185-243 == :318-384): fills beff_s (bilinear) and beff_q.
static inline void site_field(const OrcHam& H, long i, long k, const double* emomM, double* bs, double* bq) {
   bs[0] = bs[1] = bs[2] = 0.0;
   bq[0] = bq[1] = bq[2] = 0.0;
   if (H.do_jtensor != 1) heisenberg_field(H, i, k, emomM, bs);
   else tensor_field(H, i, k, emomM, bs);
   if (H.do_dm == 1) dm_field(H, i, k, emomM, bs);
   if (H.do_bq == 1) bq_field(H, i, k, emomM, bq);
   if (H.do_anisotropy == 1) {
      const double* m = emomM + 3 * ((i - 1) + (size_t)H.Natom * (k - 1));
      const int t = H.taniso[i - 1];
      if (t == 1) uniaxial_field(H, i, m, bs);
      else if (t == 2) cubic_field(H, i, m, bs);
      else if (t == 7) {
         uniaxial_field(H, i, m, bs);
         double tf[3] = {0, 0, 0};
         cubic_field(H, i, m, tf);
         for (int a = 0; a < 3; a++) bq[a] = bq[a] + tf[a] * H.sb[i - 1];
      }
   }
}

// effective_field_full (hamiltonianactions.f90:108-252).  time_external_field == 0.
// beff1/beff2 may be NULL.  Returns energy (mRy) like the reference.
double orc_effective_field(const OrcHam* Hp, const double* emomM, const double* external_field, double* beff,
                           double* beff1, double* beff2, double mub, double mry) {
   const OrcHam& H = *Hp;
   double energy = 0.0;
   const long N = H.Natom, M = H.Mensemble;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : energy)
   for (long k = 1; k <= M; k++)
      for (long i = 1; i <= N; i++) {
         double bs[3], bq[3];
         site_field(H, i, k, emomM, bs, bq);
         const size_t o = 3 * ((i - 1) + (size_t)N * (k - 1));
         double b2[3];
         for (int a = 0; a < 3; a++) b2[a] = bq[a] + external_field[o + a] + 0.0;
         if (beff1) for (int a = 0; a < 3; a++) beff1[o + a] = bs[a];
         if (beff2) for (int a = 0; a < 3; a++) beff2[o + a] = b2[a];
         for (int a = 0; a < 3; a++) beff[o + a] = 0.0 + bs[a] + b2[a];
         double tf[3];
         for (int a = 0; a < 3; a++) tf[a] = 0.5 * (bs[a] + 2.0 * bq[a] + 2.0 * external_field[o + a] + 0.0);
         energy = energy - emomM[o] * tf[0] - emomM[o + 1] * tf[1] - emomM[o + 2] * tf[2];
      }
   return energy * mub / mry;
}

// calc_energy (source/Hamiltonian/energy.f90:181-342 with update_ene :559-571, do_lsf N): per ensemble the sums over
// atoms of -factor * m_i . b_term(i), factor 1/2 for the bilinear terms and the anisotropy, 1/4 for the biquadratic
// term, 1 for the external field, then divided by the number of atoms.  terms(5, M) = exchange (the "Heis-Tens" pair
// energy when do_jtensor 1), anisotropy, DM, biquadratic, Zeeman -- in field units; totenergy.*.out prints the ensemble
// means times fcinv = mub / mry (:170, :383-398), which is what the caller applies.
void orc_energy_terms(const OrcHam* Hp, const double* emomM, const double* external_field, double* terms) {
   const OrcHam& H = *Hp;
   const long N = H.Natom, M = H.Mensemble;
   for (long k = 1; k <= M; k++) {
      double exc = 0.0, edm = 0.0, ebq = 0.0, eani = 0.0, eext = 0.0;
      for (long i = 1; i <= N; i++) {
         const double* m = emomM + 3 * ((i - 1) + (size_t)N * (k - 1));
         auto upd = [&](const double* b, double factor) { return -factor * (m[0] * b[0] + m[1] * b[1] + m[2] * b[2]); };
         double b[3] = {0.0, 0.0, 0.0};
         if (H.do_jtensor != 1) heisenberg_field(H, i, k, emomM, b);
         else tensor_field(H, i, k, emomM, b);
         exc = exc + upd(b, 0.5);
         if (H.do_jtensor != 1 && H.do_dm == 1) {
            double d[3] = {0.0, 0.0, 0.0};
            dm_field(H, i, k, emomM, d);
            edm = edm + upd(d, 0.5);
         }
         if (H.do_bq == 1) {
            double q[3] = {0.0, 0.0, 0.0};
            bq_field(H, i, k, emomM, q);
            ebq = ebq + upd(q, 0.25);
         }
         if (H.do_anisotropy == 1) {
            const int t = H.taniso[i - 1];
            double a[3] = {0.0, 0.0, 0.0};
            if (t == 1) { uniaxial_field(H, i, m, a); eani = eani + upd(a, 0.5); }
            else if (t == 2) { cubic_field(H, i, m, a); eani = eani + upd(a, 0.5); }
            else if (t == 7) {
               double c[3] = {0.0, 0.0, 0.0};
               uniaxial_field(H, i, m, a);
               cubic_field(H, i, m, c);
               for (int x = 0; x < 3; x++) a[x] = a[x] + c[x] * H.sb[i - 1];
               eani = eani + upd(a, 0.5);
            }
         }
         const double* he = external_field + 3 * ((i - 1) + (size_t)N * (k - 1));
         const double hb[3] = {0.0 + he[0], 0.0 + he[1], 0.0 + he[2]};
         eext = eext + upd(hb, 1.0);
      }
      double* t = terms + 5 * (k - 1);
      t[0] = exc / N; t[1] = eani / N; t[2] = edm / N; t[3] = ebq / N; t[4] = eext / N;
   }
}

// ---- midpoint (SDEalgh 1) ---------------------------------------------------------------------
// Cayley update shared by predictor and corrector (midpoint.f90:153-164 / :303-313): returns et*detAi.
static inline void cayley(const double* e, const double* A, double* out) {
   const double detAi = 1.0 / (1.0 + (A[0] * A[0] + A[1] * A[1] + A[2] * A[2]));
   double a2[3];
   a2[0] = e[0] + e[1] * A[2] - e[2] * A[1];
   a2[1] = e[1] + e[2] * A[0] - e[0] * A[2];
   a2[2] = e[2] + e[0] * A[1] - e[1] * A[0];
   double et[3];
   et[0] = a2[0] * (1 + A[0] * A[0]) + a2[1] * (A[0] * A[1] + A[2]) + a2[2] * (A[0] * A[2] - A[1]);
   et[1] = a2[0] * (A[1] * A[0] - A[2]) + a2[1] * (1 + A[1] * A[1]) + a2[2] * (A[1] * A[2] + A[0]);
   et[2] = a2[0] * (A[2] * A[0] + A[1]) + a2[1] * (A[2] * A[1] - A[0]) + a2[2] * (1 + A[2] * A[2]);
   out[0] = et[0] * detAi;
   out[1] = et[1] * detAi;
   out[2] = et[2] * detAi;
}

// btorque(3,N,M): the spin-transfer-torque field of the integrators (stt /= 'N': midpoint.f90:86-97 / :242-252 copy it into
// btorque_full, depondt.f90:100-113 / :255-262 add stt_fac*btorque to bdup).  A module array in the reference; here a module
// pointer set by orc_set_btorque (NULL = stt 'N': btorque_full = 0).  SHE / SOT torques are not on this path.
static const double* g_btorque = nullptr;
void orc_set_btorque(const double* bt) { g_btorque = bt; }

// smodeulermpt (midpoint.f90:34-181), Nred = Natom.
void orc_midpoint_first(int Natom, int Mensemble, const double* Landeg, double bn, const double* lambda1_array,
                        const double* beff, const double* emom, double* emom2, double* emomM, const double* mmom,
                        double deltat, const double* ranv, double* thermal_field, double gama) {
   const long N = Natom, M = Mensemble;
#pragma omp parallel for collapse(2) schedule(static)
   for (long i = 1; i <= N; i++)
      for (long j = 1; j <= M; j++) {
         const size_t o = 3 * ((i - 1) + (size_t)N * (j - 1));
         const double lam = lambda1_array[i - 1];
         const double lldamp = 1.0 / (1.0 + lam * lam);
         const double dt = deltat * bn * gama * lldamp;
         const double sqrtdt = std::sqrt(dt);
         const double dtg = dt * Landeg[i - 1];
         const double sqrtdtg = sqrtdt * Landeg[i - 1];
         const double* e = emom + o;
         const double* b = beff + o;
         const double* r = ranv + o;
         const double zero3[3] = {0.0, 0.0, 0.0};
         const double* bt = g_btorque ? g_btorque + o : zero3;      // btorque_full (midpoint.f90:86-97)
         double a1[3], s1[3], A[3];
         a1[0] = -bt[0] - b[0] - lam * (e[1] * b[2] - e[2] * b[1]);
         a1[1] = -bt[1] - b[1] - lam * (e[2] * b[0] - e[0] * b[2]);
         a1[2] = -bt[2] - b[2] - lam * (e[0] * b[1] - e[1] * b[0]);
         s1[0] = -r[0] - lam * (e[1] * r[2] - e[2] * r[1]);
         s1[1] = -r[1] - lam * (e[2] * r[0] - e[0] * r[2]);
         s1[2] = -r[2] - lam * (e[0] * r[1] - e[1] * r[0]);
         if (thermal_field) for (int a = 0; a < 3; a++) thermal_field[o + a] = s1[a];
         for (int a = 0; a < 3; a++) A[a] = 0.5 * dtg * a1[a] + 0.5 * sqrtdtg * s1[a];
         double et[3];
         cayley(e, A, et);
         for (int a = 0; a < 3; a++) et[a] = 0.5 * (e[a] + et[a]);
         const double m = mmom[(i - 1) + (size_t)N * (j - 1)];
         for (int a = 0; a < 3; a++) {
            emom2[o + a] = et[a];
            emomM[o + a] = et[a] * m;
         }
      }
}

// modeulermpf (midpoint.f90:188-323)
void orc_midpoint_second(int Natom, int Mensemble, const double* Landeg, double bn, const double* lambda1_array,
                         const double* beff, const double* emom, double* emom2, double deltat, const double* ranv,
                         double gama) {
   const long N = Natom, M = Mensemble;
#pragma omp parallel for collapse(2) schedule(static)
   for (long i = 1; i <= N; i++)
      for (long j = 1; j <= M; j++) {
         const size_t o = 3 * ((i - 1) + (size_t)N * (j - 1));
         const double lam = lambda1_array[i - 1];
         const double lldamp = 1.0 / (1.0 + lam * lam);
         const double dt = deltat * bn * gama * lldamp;
         const double sqrtdt = std::sqrt(dt);
         const double dtg = dt * Landeg[i - 1];
         const double sqrtdtg = sqrtdt * Landeg[i - 1];
         double etp[3] = {emom2[o], emom2[o + 1], emom2[o + 2]};
         const double* b = beff + o;
         const double* r = ranv + o;
         const double zero3[3] = {0.0, 0.0, 0.0};
         const double* bt = g_btorque ? g_btorque + o : zero3;      // btorque_full (midpoint.f90:242-252)
         double a1[3], s1[3], A[3];
         a1[0] = -bt[0] - b[0] - lam * (etp[1] * b[2] - etp[2] * b[1]);
         a1[1] = -bt[1] - b[1] - lam * (etp[2] * b[0] - etp[0] * b[2]);
         a1[2] = -bt[2] - b[2] - lam * (etp[0] * b[1] - etp[1] * b[0]);
         s1[0] = -r[0] - lam * (etp[1] * r[2] - etp[2] * r[1]);
         s1[1] = -r[1] - lam * (etp[2] * r[0] - etp[0] * r[2]);
         s1[2] = -r[2] - lam * (etp[0] * r[1] - etp[1] * r[0]);
         for (int a = 0; a < 3; a++) A[a] = 0.5 * dtg * a1[a] + 0.5 * sqrtdtg * s1[a];
         double out[3];
         cayley(emom + o, A, out);
         for (int a = 0; a < 3; a++) emom2[o + a] = out[a];
      }
}

// rannum scaling for llg=1 (randomnumbers.f90:667-670, 735-746): ranv holds N(0,1) on entry.
void orc_rannum_scale(int Natom, int Mensemble, const double* lambda1_array, double bn, const double* mmomi,
                      const double* Temp_array, double temprescale, double k_bolt, double gama, double mub,
                      double* ranv) {
   const long N = Natom, M = Mensemble;
   for (long j = 1; j <= M; j++)
      for (long i = 1; i <= N; i++) {
         const double lam = lambda1_array[i - 1];
         const double Dk = (lam / (1 + lam * lam) * k_bolt / gama / (mub)) * (gama / bn);
         const double D = Dk * mmomi[(i - 1) + (size_t)N * (j - 1)] * Temp_array[i - 1] * temprescale;
         const double sigma = std::sqrt(2.0 * D);
         const size_t o = 3 * ((i - 1) + (size_t)N * (j - 1));
         ranv[o] = ranv[o] * sigma;
         ranv[o + 1] = ranv[o + 1] * sigma;
         ranv[o + 2] = ranv[o + 2] * sigma;
      }
}

// ---- Depondt (SDEalgh 5) ----------------------------------------------------------------------
static inline void rodrigues(const double* bd, const double* e, double delta_t, double gama, double lldamp, double* mrod) {
   double Bnorm = bd[0] * bd[0] + bd[1] * bd[1] + bd[2] * bd[2];
   Bnorm = std::sqrt(Bnorm) + 1.0e-15;
   const double hx = bd[0] / Bnorm, hy = bd[1] / Bnorm, hz = bd[2] / Bnorm;
   const double v = Bnorm * delta_t * gama * lldamp;
   const double cosv = std::cos(v), sinv = std::sin(v);
   const double u = 1.0 - cosv;
   mrod[0] = hx * hx * u * e[0] + cosv * e[0] + hx * hy * u * e[1] - hz * sinv * e[1] + hx * hz * u * e[2] + hy * sinv * e[2];
   mrod[1] = hy * hx * u * e[0] + hz * sinv * e[0] + hy * hy * u * e[1] + cosv * e[1] + hy * hz * u * e[2] - hx * sinv * e[2];
   mrod[2] = hx * hz * u * e[0] - hy * sinv * e[0] + hz * hy * u * e[1] + hx * sinv * e[1] + hz * hz * u * e[2] + cosv * e[2];
}

// depondt_evolve_first (depondt.f90:50-193).  btherm holds N(0,1) on entry and sigma-scaled noise on exit
// (the module array the corrector re-reads).
void orc_depondt_first(int Natom, int Mensemble, const double* lambda1_array, const double* beff, double* b2eff,
                       double* emom, double* emom2, double* emomM, const double* mmom, double delta_t,
                       const double* Temp_array, double temprescale, double* btherm, double k_bolt, double gama,
                       double mub) {
   const long N = Natom, M = Mensemble;
#pragma omp parallel for collapse(2) schedule(static)
   for (long k = 1; k <= M; k++)
      for (long i = 1; i <= N; i++) {
         const size_t o = 3 * ((i - 1) + (size_t)N * (k - 1));
         const double lam = lambda1_array[i - 1];
         const double m = mmom[(i - 1) + (size_t)N * (k - 1)];
         const double Dp = (2.0 * lam * k_bolt) / (delta_t * gama * mub);
         const double sigma = std::sqrt(Dp * temprescale * Temp_array[i - 1] / m);
         double bloc[3], bdup[3] = {0.0, 0.0, 0.0};
         if (g_btorque) for (int a = 0; a < 3; a++) bdup[a] = bdup[a] + 1.0 * g_btorque[o + a];   // bdup = 0 + stt_fac*btorque (depondt.f90:100-113, :255-262)
         for (int a = 0; a < 3; a++) {
            btherm[o + a] = btherm[o + a] * sigma;
            bloc[a] = beff[o + a] + btherm[o + a];
         }
         double* e = emom + o;
         bdup[0] = bdup[0] + bloc[0] + lam * e[1] * bloc[2] - lam * e[2] * bloc[1];
         bdup[1] = bdup[1] + bloc[1] + lam * e[2] * bloc[0] - lam * e[0] * bloc[2];
         bdup[2] = bdup[2] + bloc[2] + lam * e[0] * bloc[1] - lam * e[1] * bloc[0];
         const double lldamp = 1.0 / (1.0 + lam * lam);
         double mrod[3];
         rodrigues(bdup, e, delta_t, gama, lldamp, mrod);
         for (int a = 0; a < 3; a++) {
            emom2[o + a] = e[a];
            emomM[o + a] = mrod[a] * m;
         }
         for (int a = 0; a < 3; a++) {
            e[a] = mrod[a];
            b2eff[o + a] = bdup[a];
         }
      }
}

// depondt_evolve_second (depondt.f90:203-333)
void orc_depondt_second(int Natom, int Mensemble, const double* lambda1_array, const double* beff,
                        const double* b2eff, double* emom, double* emom2, double delta_t, const double* btherm,
                        double gama) {
   const long N = Natom, M = Mensemble;
#pragma omp parallel for collapse(2) schedule(static)
   for (long k = 1; k <= M; k++)
      for (long i = 1; i <= N; i++) {
         const size_t o = 3 * ((i - 1) + (size_t)N * (k - 1));
         const double lam = lambda1_array[i - 1];
         double bloc[3], bdup[3] = {0.0, 0.0, 0.0};
         if (g_btorque) for (int a = 0; a < 3; a++) bdup[a] = bdup[a] + 1.0 * g_btorque[o + a];   // bdup = 0 + stt_fac*btorque (depondt.f90:100-113, :255-262)
         for (int a = 0; a < 3; a++) bloc[a] = beff[o + a] + btherm[o + a];
         double* e = emom + o;
         bdup[0] = bdup[0] + bloc[0] + lam * e[1] * bloc[2] - lam * e[2] * bloc[1];
         bdup[1] = bdup[1] + bloc[1] + lam * e[2] * bloc[0] - lam * e[0] * bloc[2];
         bdup[2] = bdup[2] + bloc[2] + lam * e[0] * bloc[1] - lam * e[1] * bloc[0];
         for (int a = 0; a < 3; a++) bdup[a] = 0.5 * bdup[a] + 0.5 * b2eff[o + a];
         for (int a = 0; a < 3; a++) e[a] = emom2[o + a];
         const double lldamp = 1.0 / (1.0 + lam * lam);
         double mrod[3];
         rodrigues(bdup, e, delta_t, gama, lldamp, mrod);
         for (int a = 0; a < 3; a++) emom2[o + a] = mrod[a];
      }
}

// moment_update = calcm + copym (updatemoments.f90:19-145), initexc /= 'I'.
void orc_moment_update(int Natom, int Mensemble, double* mmom, const double* mmom0, double* mmom2, double* emom,
                       const double* emom2, double* emomM, double* mmomi, int mompar) {
   const long N = Natom, M = Mensemble;
   for (long j = 0; j < M; j++)
      for (long i = 0; i < N; i++) {
         const size_t q = i + (size_t)N * j;
         if (mompar == 1) { mmom2[q] = std::fmax(mmom0[q] * std::fabs(emom2[3 * q + 2]), 1e-4); mmom[q] = mmom2[q]; }
         else if (mompar == 2) { mmom2[q] = std::fmax(mmom0[q] * (emom2[3 * q + 2] * emom2[3 * q + 2]), 1.0e-4); mmom[q] = mmom2[q]; }
         else mmom2[q] = mmom[q];
      }
#pragma omp parallel for collapse(2) schedule(static)
   for (long j = 0; j < M; j++)
      for (long i = 0; i < N; i++) {
         const size_t q = i + (size_t)N * j;
         for (int a = 0; a < 3; a++) {
            emom[3 * q + a] = emom2[3 * q + a];
            emomM[3 * q + a] = emom2[3 * q + a] * mmom2[q];
         }
         mmomi[q] = 1.0 / mmom[q];
      }
}

// ---- fused drivers: one reference time step (sd_driver.f90:668-764) ------------------------------
// SDEalgh 1 or 5; noise: gauss (3,N,M) N(0,1) per step supplied by the caller (NULL => T=0 / zeros).
// work must hold 4 arrays of 3*N*M doubles (beff, b2eff, ranv/btherm, emom2) + 2*N*M (mmom2, mmomi).
void orc_sd_step(const OrcHam* H, int SDEalgh, double* emom, double* emomM, double* mmom, const double* mmom0,
                 const double* external_field, const double* Landeg, const double* lambda1_array,
                 const double* Temp_array, double temprescale, double delta_t, int mompar, const double* gauss,
                 double gama, double k_bolt, double mub, double mry, double* work, const unsigned char* frozen) {
   // frozen (may be NULL): frozen[i-1] != 0 for atoms that are NOT in red_atom_list (Nred < Natom).  The reference's loops
   // (midpoint.f90:123, depondt.f90:138 and their second halves) simply never visit such an atom: emom2 keeps the value
   // magninit gave it (emom2 = emom, magnetizationinit.f90:538), copym writes that back.  Restated here by equivalence: the
   // stage routines run over every atom and the rows of the frozen atoms are put back after each stage.
   const size_t NM = (size_t)H->Natom * H->Mensemble;
   std::vector<double> keep;
   if (frozen) keep.assign(emom, emom + 3 * NM);
   auto put_back = [&](double* emom2_) {
      if (!frozen) return;
      for (long k = 0; k < H->Mensemble; k++)
         for (long i = 0; i < H->Natom; i++)
            if (frozen[i]) {
               const size_t q = (size_t)i + (size_t)H->Natom * k;
               for (int a = 0; a < 3; a++) {
                  emom[3 * q + a] = keep[3 * q + a];
                  emom2_[3 * q + a] = keep[3 * q + a];
                  emomM[3 * q + a] = keep[3 * q + a] * mmom[q];
               }
            }
   };
   double* beff = work;
   double* b2eff = work + 3 * NM;
   double* ranv = work + 6 * NM;
   double* emom2 = work + 9 * NM;
   double* mmom2 = work + 12 * NM;
   double* mmomi = work + 13 * NM;
   const double bn = 1.0;
   for (size_t q = 0; q < NM; q++) mmomi[q] = 1.0 / mmom[q];
   if (gauss) std::memcpy(ranv, gauss, 3 * NM * sizeof(double)); else std::memset(ranv, 0, 3 * NM * sizeof(double));
   orc_effective_field(H, emomM, external_field, beff, nullptr, nullptr, mub, mry);
   if (SDEalgh == 1) {
      if (gauss) orc_rannum_scale(H->Natom, H->Mensemble, lambda1_array, bn, mmomi, Temp_array, temprescale, k_bolt, gama, mub, ranv);
      orc_midpoint_first(H->Natom, H->Mensemble, Landeg, bn, lambda1_array, beff, emom, emom2, emomM, mmom, delta_t, ranv, nullptr, gama);
      put_back(emom2);
      orc_effective_field(H, emomM, external_field, beff, nullptr, nullptr, mub, mry);
      orc_midpoint_second(H->Natom, H->Mensemble, Landeg, bn, lambda1_array, beff, emom, emom2, delta_t, ranv, gama);
      put_back(emom2);
   } else {
      orc_depondt_first(H->Natom, H->Mensemble, lambda1_array, beff, b2eff, emom, emom2, emomM, mmom, delta_t, Temp_array, temprescale, ranv, k_bolt, gama, mub);
      put_back(emom2);
      orc_effective_field(H, emomM, external_field, beff, nullptr, nullptr, mub, mry);
      orc_depondt_second(H->Natom, H->Mensemble, lambda1_array, beff, b2eff, emom, emom2, delta_t, ranv, gama);
      put_back(emom2);
   }
   orc_moment_update(H->Natom, H->Mensemble, mmom, mmom0, mmom2, emom, emom2, emomM, mmomi, mompar);
}

// ---- observables ------------------------------------------------------------------------------
// buffer_avrg (prn_averages.f90:437-447): m(:,k) = sum_i emomM(:,i,k), atom-major accumulation order.
void orc_sum_moments(int Natom, int Mensemble, const double* emomM, double* m /*(3,M)*/) {
   for (long k = 0; k < Mensemble; k++) m[3 * k] = m[3 * k + 1] = m[3 * k + 2] = 0.0;
   for (long i = 0; i < Natom; i++)
      for (long k = 0; k < Mensemble; k++)
         for (int a = 0; a < 3; a++) m[3 * k + a] = m[3 * k + a] + emomM[a + 3 * (i + (size_t)Natom * k)];
}

// ---- Monte Carlo -------------------------------------------------------------------------------
// The current-state DM term of calculate_energy mixes emom and emomM (montecarlo_common.f90:611-616); identical to the
// consistent form for |m| = 1.  orc_set_dm_energy_quirk(0) selects the consistent form -m_i . (m_j x D) with emomM only, which
// is what the product computes (documented deviation): used by the deterministic chain-parity tests on systems
// with |m| /= 1.  Default 1 = the reference as written.
static int g_dm_quirk = 1;
void orc_set_dm_energy_quirk(int on) { g_dm_quirk = on; }

// calculate_energy (montecarlo_common.f90:431-865) for exchange(+DM+BQ+anisotropy+Zeeman), returns de.
// Reference quirk kept: the DM current-state term mixes emom and emomM (:611-616).
static double mc_delta_e(const OrcHam& H, const double* emomM, const double* emom, const double* mmom, long iflip,
                         const double* newmom, const double* extfield, long k, double mub) {
   const long N = H.Natom;
   const double* eM = emomM + 3 * (size_t)N * (k - 1);
   const double* eU = emom + 3 * (size_t)N * (k - 1);
   double e_c = 0.0, e_t = 0.0;
   double trial[3];
   const double mm = mmom[(iflip - 1) + (size_t)N * (k - 1)];
   for (int a = 0; a < 3; a++) trial[a] = newmom[a] * mm;
   const int ih = H.aHam[iflip - 1];
   const double* mi = eM + 3 * (iflip - 1);
   for (int j = 1; j <= H.nlistsize[ih - 1]; j++) {
      const double c = H.ncoup[(j - 1) + (size_t)H.max_no_neigh * (ih - 1)];
      const double* mj = eM + 3 * ((long)H.nlist[(j - 1) + (size_t)H.max_no_neigh * (iflip - 1)] - 1);
      e_c = e_c - c * (mi[0] * mj[0] + mi[1] * mj[1] + mi[2] * mj[2]);
      e_t = e_t - c * (trial[0] * mj[0] + trial[1] * mj[1] + trial[2] * mj[2]);
   }
   if (H.do_anisotropy == 1) {
      const int t = H.taniso[iflip - 1];
      const double* ea = H.eaniso + 3 * (iflip - 1);
      const double* kk = H.kaniso + 2 * (iflip - 1);
      // x**2*y**2 + ... : each square is formed first, as the Fortran '**2' does
      auto cub = [](const double* v) { return (v[0] * v[0]) * (v[1] * v[1]) + (v[1] * v[1]) * (v[2] * v[2]) + (v[2] * v[2]) * (v[0] * v[0]); };
      auto cub6 = [](const double* v) { return (v[0] * v[0]) * (v[1] * v[1]) * (v[2] * v[2]); };
      auto p4 = [](double t) { double t2 = t * t; return t2 * t2; };
      if (t == 1) {
         const double tta = mi[0] * ea[0] + mi[1] * ea[1] + mi[2] * ea[2];
         const double ttb = trial[0] * ea[0] + trial[1] * ea[1] + trial[2] * ea[2];
         e_c = e_c + kk[0] * (tta * tta) + kk[1] * p4(tta);
         e_t = e_t + kk[0] * (ttb * ttb) + kk[1] * p4(ttb);
      } else if (t == 2) {
         e_c = e_c - kk[0] * cub(mi) - kk[1] * cub6(mi);
         e_t = e_t - kk[0] * cub(trial) - kk[1] * cub6(trial);
      }
      if (t == 7) {
         const double tta = mi[0] * ea[0] + mi[1] * ea[1] + mi[2] * ea[2];
         const double ttb = trial[0] * ea[0] + trial[1] * ea[1] + trial[2] * ea[2];
         e_c = e_c + kk[0] * (tta * tta) + kk[1] * p4(tta);
         e_t = e_t + kk[0] * (ttb * ttb) + kk[1] * p4(ttb);
         const double aw1 = kk[0] * H.sb[iflip - 1], aw2 = kk[1] * H.sb[iflip - 1];
         e_c = e_c + aw1 * cub(mi) + aw2 * cub6(mi);
         e_t = e_t + aw1 * cub(trial) + aw2 * cub6(trial);
      }
   }
   if (H.do_dm == 1) {
      const double* ui = g_dm_quirk ? eU + 3 * (iflip - 1) : mi;
      for (int j = 1; j <= H.dmlistsize[ih - 1]; j++) {
         const double* D = H.dm_vect + 3 * ((j - 1) + (size_t)H.max_no_dmneigh * (ih - 1));
         const double* mj = eM + 3 * ((long)H.dmlist[(j - 1) + (size_t)H.max_no_dmneigh * (iflip - 1)] - 1);
         e_c = e_c - D[0] * (mi[1] * mj[2] - ui[2] * mj[1]) - D[1] * (mi[2] * mj[0] - mi[0] * mj[2]) -
               D[2] * (ui[0] * mj[1] - mi[1] * mj[0]);
         e_t = e_t - D[0] * (trial[1] * mj[2] - trial[2] * mj[1]) - D[1] * (trial[2] * mj[0] - trial[0] * mj[2]) -
               D[2] * (trial[0] * mj[1] - trial[1] * mj[0]);
      }
   }
   if (H.do_bq == 1) {
      for (int j = 1; j <= H.bqlistsize[ih - 1]; j++) {
         const double c = H.j_bq[(j - 1) + (size_t)H.nn_bq_tot * (ih - 1)];
         const double* mj = eM + 3 * ((long)H.bqlist[(j - 1) + (size_t)H.nn_bq_tot * (iflip - 1)] - 1);
         double d = mj[0] * mi[0] + mj[1] * mi[1] + mj[2] * mi[2];
         e_c = e_c - c * (d * d);
         d = mj[0] * trial[0] + mj[1] * trial[1] + mj[2] * trial[2];
         e_t = e_t - c * (d * d);
      }
   }
   e_c = e_c - extfield[0] * mi[0] - extfield[1] * mi[1] - extfield[2] * mi[2];
   e_t = e_t - extfield[0] * trial[0] - extfield[1] * trial[1] - extfield[2] * trial[2];
   return mub * (e_t - e_c);
}

// choose_random_flip (montecarlo_common.f90:25-79)
static void choose_random_flip(const double* e, double* newmom, double delta, const double* rn, const double* gn) {
   const double pi = 3.141592653589793;
   const int ftype = (int)std::floor(3 * rn[0]);
   if (ftype == 0) {
      const double phi = rn[1] * 2 * pi;
      const double theta = std::acos(1 - 2 * rn[2]);
      newmom[0] = std::sin(theta) * std::cos(phi);
      newmom[1] = std::sin(theta) * std::sin(phi);
      newmom[2] = std::cos(theta);
   } else if (ftype == 1) {
      double g[3] = {gn[0] * delta, gn[1] * delta, gn[2] * delta};
      const double l = std::sqrt((e[0] + g[0]) * (e[0] + g[0]) + (e[1] + g[1]) * (e[1] + g[1]) + (e[2] + g[2]) * (e[2] + g[2]));
      for (int a = 0; a < 3; a++) newmom[a] = (e[a] + g[a]) / l;
   } else {
      for (int a = 0; a < 3; a++) newmom[a] = -e[a];
   }
}

// mc_evolve (montecarlo.f90:44-273), modes 'M' and 'H', sequential sweep (OMP_NUM_THREADS=1 semantics).
// Random inputs are supplied by the caller in the reference's draw order:
//   flipprob_m (3,N,M) uniform, flipprob_g (3,N,M) N(0,1), mflip (N,M) uniform [H only], flipprob_a (N,M) uniform.
// The heat bath reads its Zeeman field from external_field(3,N,M) (reference quirk, montecarlo.f90:231-237).
void orc_mc_sweep(const OrcHam* Hp, char mode, const int* iflip_a, double* emomM, double* emom, const double* mmom,
                  const double* extfield, const double* external_field, double temperature, double temprescale,
                  const double* flipprob_m, const double* flipprob_g, const double* mflip, const double* flipprob_a,
                  double k_bolt, double mub) {
   const OrcHam& H = *Hp;
   const long N = H.Natom, M = H.Mensemble;
   const double pi = 3.141592653589793;
   const double dbl_tolerance = (double)1e-14f;
   // delta=(2.0/25.0)*(k_bolt*temperature/mub)**(0.20_dblprec): 2.0/25.0 is a default-real constant expression
   const double delta = (double)(2.0f / 25.0f) * std::pow(k_bolt * temperature / mub, 0.20);
   std::vector<double> newmom_a((size_t)3 * N * M);
   for (long i = 1; i <= N; i++)
      for (long k = 1; k <= M; k++) {
         const size_t o = 3 * ((i - 1) + (size_t)N * (k - 1));
         choose_random_flip(emom + o, newmom_a.data() + o, delta, flipprob_m + o, flipprob_g + o);
      }
   for (long i = 1; i <= N; i++)
      for (long k = 1; k <= M; k++) {
         const long ia = iflip_a[i - 1];
         const size_t o = 3 * ((ia - 1) + (size_t)N * (k - 1));
         if (mode == 'H') {
            double bs[3], bq[3], tot[3];
            site_field(H, ia, k, emomM, bs, bq);
            for (int a = 0; a < 3; a++) tot[a] = 0.0 + bs[a] + (bq[a] + external_field[o + a] + 0.0);
            // flip_h (montecarlo_common.f90:371-422)
            const double mm = mmom[(ia - 1) + (size_t)N * (k - 1)];
            const double beta = 1.0 / k_bolt / (temprescale * temperature);
            double zfc[3];
            for (int a = 0; a < 3; a++) zfc[a] = beta * tot[a] * mub * mm;
            const double zarg = std::sqrt(zfc[0] * zfc[0] + zfc[1] * zfc[1] + zfc[2] * zfc[2]);
            const double zctheta = zfc[2] / zarg;
            const double zstheta = std::sqrt(1.0 - zctheta * zctheta) + dbl_tolerance;
            const double zcphi = zfc[0] / (zarg * zstheta);
            const double zsphi = zfc[1] / (zarg * zstheta);
            const double q = mflip[(i - 1) + (size_t)N * (k - 1)];
            const double ctheta = 1.0 + (1.0 / zarg) * std::log((1.0 - std::exp(-2.0 * zarg)) * q + std::exp(-2.0 * zarg) + dbl_tolerance);
            const double stheta = std::sqrt(1.0 - ctheta * ctheta);
            const double phi = pi * (2.0 * flipprob_a[(i - 1) + (size_t)N * (k - 1)] - 1.0);
            const double st[3] = {stheta * std::cos(phi), stheta * std::sin(phi), ctheta};
            emom[o] = zcphi * zctheta * st[0] - zsphi * st[1] + zcphi * zstheta * st[2];
            emom[o + 1] = zsphi * zctheta * st[0] + zcphi * st[1] + zsphi * zstheta * st[2];
            emom[o + 2] = -zstheta * st[0] + zctheta * st[2];
            for (int a = 0; a < 3; a++) emomM[o + a] = mm * emom[o + a];
         } else {
            const double de = mc_delta_e(H, emomM, emom, mmom, ia, newmom_a.data() + o, extfield, k, mub);
            // flip_a (montecarlo_common.f90:190-200)
            const double beta = 1.0 / k_bolt / (temprescale * temperature + 1.0e-15);
            if (de <= 0.0 || flipprob_a[(i - 1) + (size_t)N * (k - 1)] < std::exp(-beta * de)) {
               const double mm = mmom[(ia - 1) + (size_t)N * (k - 1)];
               for (int a = 0; a < 3; a++) {
                  emom[o + a] = newmom_a[o + a];
                  emomM[o + a] = mm * newmom_a[o + a];
               }
            }
         }
      }
}

}  // extern "C"
