/* The loop body of sd_minimal (source/sd_driver.f90:1162-1326) behind the explicit C ABI, from plain C: what a host written
 * in C (or Fortran through bind(c)) does with the tables it already owns.  A two-atom dimer with
 * one neighbour each precesses for 100 midpoint steps (asd_sd_run: samples of sum M stay on the device until the call ends).  Build:  gcc -std=c99 -I include examples/sd_minimal.c \
 *   -L uppasd_b200 -luppasd_b200 -Wl,-rpath,$PWD/uppasd_b200 -o sd_minimal   (needs a CUDA device at run time). */
#include <stdio.h>
#include "uppasd_b200.h"

int main(void) {
   asd_engine* eng = NULL;
   if (asd_create(&eng, -1)) { fprintf(stderr, "%s\n", asd_last_error()); return 1; }   /* fails loudly without a GPU */
   const int N = 2, M = 1, nHam = 2;
   const int aHam[2] = {1, 2};
   const int nlist[2] = {2, 1}, nlistsize[2] = {1, 1};          /* nlist(1,N): each atom's only neighbour */
   const double ncoup[2] = {10.0, 10.0};                        /* ncoup(1,nHam), field units */
   const double emom[6] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};       /* emom(3,N,M) */
   const double mmom[2] = {1.0, 1.0};
   const double landeg[2] = {1.0, 1.0}, lambda1[2] = {0.1, 0.1}, temp[2] = {0.0, 0.0};
   double out[6], outM[6], mm[2], msum[3], samples[3 * 10];
   long nsamples = 0;
   int rc = 0;
   rc |= asd_set_constants(eng, 1.760859644e11, 1.38064852e-23, 9.274009994e-24, 2.179872325e-21);
   rc |= asd_set_system(eng, N, M, nHam, aHam);
   rc |= asd_set_exchange(eng, 1, nlist, nlistsize, ncoup);
   rc |= asd_set_llg(eng, 1, 1.0e-16, landeg, lambda1, temp, 1.0, 0, 1ULL);
   rc |= asd_set_moments(eng, emom, mmom, mmom);
   rc |= asd_commit(eng);
   /* the measurement-phase loop in one call: 100 steps, sum M sampled on the device after every 10th, one copy at the end */
   rc |= asd_sd_run(eng, 100, 1, 10, samples, &nsamples);
   rc |= asd_get_moments(eng, out, outM, mm);
   rc |= asd_measure(eng, msum, NULL);
   if (rc) { fprintf(stderr, "%s\n", asd_last_error()); asd_destroy(eng); return 1; }
   printf("e1 = (%.12f, %.12f, %.12f)  sum M = (%.12f, %.12f, %.12f)\n", out[0], out[1], out[2], msum[0], msum[1], msum[2]);
   printf("samples = %ld, last sample = (%.12f, %.12f, %.12f)\n", nsamples, samples[27], samples[28], samples[29]);
   asd_destroy(eng);
   return 0;
}
