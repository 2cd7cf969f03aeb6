// TEST INFRASTRUCTURE ONLY -- restatement of the reference's random-number generators.
// Nothing under uppasd_b200/ may include, link or call this file.
//
//   mtprng            source/RNG/mtprng.f90:118-138 (init), :191-245 (rand64), :248-331 (rand, real1, real2)
//   ziggurat          source/RNG/randomnumbers.f90:330-438 (r4_nor), :441-509 (r4_nor_setup)
//   fill_rngarray     source/RNG/randomnumbers.f90:548-565
//   visiting order    source/MonteCarlo/montecarlo.f90:277-301 (choose_random_atom_x)
//   initmag 1         source/System/magnetizationinit.f90:141-178
//
// The reference generator is NOT the standard MT19937: its state words are 64-bit integers and the
// twist/tempering masks are sign-extended negatives (mtprng.f90:199-207), which leaks bits 32..42 of a
// state word into bits 21..31 of the output through the unmasked first tempering shift.  This file
// reproduces that arithmetic literally with two's-complement 64-bit integers.
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace {
const int N = 624, M = 397;
struct mtprng_state {
   int mti = -1;
   int64_t mt[624];
};

inline int64_t ishft(int64_t v, int s) {  // Fortran ISHFT: logical shift, s<0 => right
   uint64_t u = (uint64_t)v;
   if (s >= 0) return (int64_t)(s >= 64 ? 0 : u << s);
   return (int64_t)(-s >= 64 ? 0 : u >> (-s));
}

void mtprng_init(int32_t seed, mtprng_state& st) {
   st.mt[0] = seed;
   for (int i = 1; i <= N - 1; i++) {
      int64_t x = st.mt[i - 1] ^ ishft(st.mt[i - 1], -30);
      // 1812433253 * x + i in wrapping 64-bit arithmetic, then masked to 32 bits
      uint64_t p = (uint64_t)1812433253LL * (uint64_t)x + (uint64_t)i;
      st.mt[i] = (int64_t)(p & 4294967295ULL);
   }
   st.mti = N;
}

int64_t mtprng_rand64(mtprng_state& st) {
   static const int64_t mag01[2] = {0, -1727483681LL};
   const int64_t UPPER_MASK = 2147483648LL, LOWER_MASK = 2147483647LL;
   const int64_t TEMPERING_B = -1658038656LL, TEMPERING_C = -272236544LL;
   int64_t r;
   if (st.mti >= N) {
      if (st.mti == -1) mtprng_init(4357, st);
      int kk;
      for (kk = 0; kk <= N - M - 1; kk++) {
         r = (st.mt[kk] & UPPER_MASK) | (st.mt[kk + 1] & LOWER_MASK);
         st.mt[kk] = (st.mt[kk + M] ^ ishft(r, -1)) ^ mag01[r & 1];
      }
      for (kk = N - M; kk <= N - 2; kk++) {
         r = (st.mt[kk] & UPPER_MASK) | (st.mt[kk + 1] & LOWER_MASK);
         st.mt[kk] = (st.mt[kk + (M - N)] ^ ishft(r, -1)) ^ mag01[r & 1];
      }
      r = (st.mt[N - 1] & UPPER_MASK) | (st.mt[0] & LOWER_MASK);
      st.mt[N - 1] = (st.mt[M - 1] ^ ishft(r, -1)) ^ mag01[r & 1];
      st.mti = 0;
   }
   r = st.mt[st.mti];
   st.mti = st.mti + 1;
   r = r ^ ishft(r, -11);
   r = 4294967295LL & (r ^ (ishft(r, 7) & TEMPERING_B));
   r = 4294967295LL & (r ^ (ishft(r, 15) & TEMPERING_C));
   r = r ^ ishft(r, -18);
   return r;
}

int32_t mtprng_rand(mtprng_state& st) {
   int64_t x = mtprng_rand64(st);
   if (x > 2147483647LL) return (int32_t)(x - 4294967296LL);
   return (int32_t)x;
}
double mtprng_rand_real1(mtprng_state& st) { return (double)mtprng_rand64(st) * (1.0 / 4294967295.0); }
double mtprng_rand_real2(mtprng_state& st) { return (double)mtprng_rand64(st) * (1.0 / 4294967296.0); }

mtprng_state state_c, state_z;
int32_t kn[128];
double fn[128], wn[128];
const double r_zig = 3.442620;

double r4_nor() {
   int32_t hz = mtprng_rand(state_z);
   int32_t iz = hz & 127;
   double value, x, y;
   // abs(hz) for hz = -2^31 overflows in both languages; unreachable in practice for the compare below
   if (std::llabs((long long)hz) < kn[iz]) {
      value = (double)hz * wn[iz];
   } else {
      for (;;) {
         if (iz == 0) {
            for (;;) {
               // log(real(u)) : the argument is converted to default (single) real first
               x = -0.2904764 * (double)std::log((float)mtprng_rand_real1(state_z));
               y = -(double)std::log((float)mtprng_rand_real1(state_z));
               if (x * x <= y + y) break;
            }
            value = (hz <= 0) ? (-r_zig - x) : (r_zig + x);
            break;
         }
         x = (double)hz * wn[iz];
         if (fn[iz] + mtprng_rand_real1(state_z) * (fn[iz - 1] - fn[iz]) < std::exp(-0.5 * x * x)) {
            value = x;
            break;
         }
         hz = mtprng_rand(state_z);
         iz = hz & 127;
         if (std::llabs((long long)hz) < kn[iz]) {
            value = (double)hz * wn[iz];
            break;
         }
      }
   }
   return value;
}
}  // namespace

extern "C" {

void orc_rng_init(int seed) { mtprng_init(seed, state_c); }           // randomnumbers.f90:109-125
void orc_rng_uniform(double* out, long len) {                         // randomnumbers.f90:128-147
   for (long i = 0; i < len; i++) out[i] = mtprng_rand_real2(state_c);
}
unsigned orc_rng_raw32() { return (unsigned)(mtprng_rand64(state_c) & 0xffffffffULL); }

void orc_zig_setup(int inseed) {                                      // randomnumbers.f90:441-509
   const double m1 = 2147483648.0, vn = 9.91256303526217e-03;
   mtprng_init(inseed, state_z);
   double dn = 3.442619855899, tn = 3.442619855899;
   double q = vn / std::exp(-0.5 * dn * dn);
   kn[0] = (int32_t)((dn / q) * m1);
   kn[1] = 0;
   wn[0] = q / m1;
   wn[127] = dn / m1;
   fn[0] = 1.0;
   fn[127] = std::exp(-0.5 * dn * dn);
   for (int i = 127; i >= 2; i--) {
      dn = std::sqrt(-2.0 * std::log(vn / dn + std::exp(-0.5 * dn * dn)));
      kn[i] = (int32_t)((dn / tn) * m1);   // kn(i+1)
      tn = dn;
      fn[i - 1] = std::exp(-0.5 * dn * dn);  // fn(i)
      wn[i - 1] = dn / m1;                   // wn(i)
   }
}
void orc_fill_rngarray(double* out, long len) {                       // randomnumbers.f90:548-565
   for (long i = 0; i < len; i++) out[i] = r4_nor();
}

// choose_random_atom_x (montecarlo.f90:277-301)
void orc_choose_random_atom_x(int Natom, int* iflip_a) {
   for (int i = 1; i <= Natom; i++) iflip_a[i - 1] = i;
   double* dshift = (double*)std::malloc(sizeof(double) * Natom);
   orc_rng_uniform(dshift, Natom);
   for (int i = 1; i <= Natom; i++) {
      int ishift = (int)(dshift[i - 1] * Natom);
      int itmp = iflip_a[i - 1];
      iflip_a[i - 1] = iflip_a[ishift % Natom];
      iflip_a[ishift % Natom] = itmp;
   }
   std::free(dshift);
}

// initmag 1 (magnetizationinit.f90:141-178), do_ralloy = 0: random directions from state_c.
void orc_initmag1(int Natom, int Mensemble, int NA, int N1, int N2, int N3, const double* mmom, double* emom,
                  double* emomM) {
   for (int I3 = 0; I3 < N3; I3++)
      for (int I2 = 0; I2 < N2; I2++)
         for (int I1 = 0; I1 < N1; I1++)
            for (int I0 = 1; I0 <= NA; I0++) {
               long i = I0 + (long)I1 * NA + (long)I2 * N1 * NA + (long)I3 * N2 * N1 * NA;
               double rn[3];
               orc_rng_uniform(rn, 3);
               double x = 2.0 * (rn[0] - 0.50), y = 2.0 * (rn[1] - 0.50), z = 2.0 * (rn[2] - 0.50);
               while (x * x + y * y + z * z > 1) {
                  orc_rng_uniform(rn, 3);
                  x = 1.0 * (rn[0] - 0.50); y = 1.0 * (rn[1] - 0.50); z = 1.0 * (rn[2] - 0.50);
               }
               double nrm = std::sqrt(x * x + y * y + z * z);
               double u[3] = {x / nrm, y / nrm, z / nrm};
               for (int j = 0; j < Mensemble; j++) {
                  double mm = mmom[(I0 - 1) + (size_t)Natom * j];
                  for (int a = 0; a < 3; a++) {
                     emom[a + 3 * ((i - 1) + (size_t)Natom * j)] = u[a];
                     emomM[a + 3 * ((i - 1) + (size_t)Natom * j)] = u[a] * mm;
                  }
               }
            }
}
}
