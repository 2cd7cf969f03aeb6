// On-device construction of the neighbour tables of a supercell from its unit-cell stencil.
//
// Restates the second half of setup_nm (source/Hamiltonian/neighbourmap.f90:248-321: translate the
// first-cell stencil nm_cell/nm_trunk to every cell, wrap periodic directions with mod(j+1000*N,N), drop
// neighbours outside open boundaries) fused with the list compaction of setup_neighbour_hamiltonian
// (source/Hamiltonian/hamiltonianinit.f90:1040-1091: keep existing neighbours in stencil order, skip an atom
// that is already in the list unless map_multiple), writing straight into the device layout.  The host
// never materialises an O(N z) table -- required for the 134 M-spin slab case.
#pragma once
#include <cuda_runtime.h>
#include "asd_device.cuh"

namespace asd {

struct LatticeDesc {
   int NA, N1, N2, N3;
   int periodic[3];
   int reduced;      // 1: device order = basis-atom major (one group per ham row), 0: original order
   int Ncell;        // N1*N2*N3
   int Ncell_pad;    // Ncell rounded up to 32
   int N, Npad;
};

__device__ __host__ __forceinline__ int lattice_slot(const LatticeDesc& d, int i0, int cell) {
   return d.reduced ? i0 * d.Ncell_pad + cell : cell * d.NA + i0;
}

// orig[] / ham[] of every device slot
__global__ void lattice_index_kernel(const LatticeDesc d, int* __restrict__ orig, int* __restrict__ ham) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= d.Npad) return;
   int o = -1, h = -1;
   if (d.reduced) {
      const int i0 = s / d.Ncell_pad, cell = s - i0 * d.Ncell_pad;
      if (cell < d.Ncell) { o = cell * d.NA + i0; h = i0; }
   } else if (s < d.N) { o = s; h = 0; }
   orig[s] = o;
   ham[s] = h;
}

// One thread per atom.  nl[z][Npad] (device slots, self beyond the list), count[Npad] accepted entries,
// cp[ncomp][z][Npad] per-atom couplings (non-reduced only, zero beyond the list).
__global__ void __launch_bounds__(128)
lattice_table_kernel(const LatticeDesc d, int maxslot, int z, int ncomp, int dedup,
                     const int* __restrict__ nslot, const int* __restrict__ cell_atom,
                     const int* __restrict__ cell_shift, const double* __restrict__ coupling,
                     int* __restrict__ nl, int* __restrict__ count, double* __restrict__ cp) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= d.Npad) return;
   int i0, cell;
   bool real;
   if (d.reduced) { i0 = s / d.Ncell_pad; cell = s - i0 * d.Ncell_pad; real = cell < d.Ncell; }
   else { real = s < d.N; cell = s / d.NA; i0 = s - cell * d.NA; }
   int n = 0;
   if (real) {
      const int ix = cell % d.N1, iy = (cell / d.N1) % d.N2, iz = cell / (d.N1 * d.N2);
      const int ns = nslot[i0];
      for (int q = 0; q < ns; q++) {
         const int j0 = cell_atom[i0 * maxslot + q] - 1;
         const int* sh = cell_shift + 3 * (i0 * maxslot + q);
         int jx = sh[0] + ix, jy = sh[1] + iy, jz = sh[2] + iz;
         if (d.periodic[0]) jx = (jx + 1000 * d.N1) % d.N1;
         if (d.periodic[1]) jy = (jy + 1000 * d.N2) % d.N2;
         if (d.periodic[2]) jz = (jz + 1000 * d.N3) % d.N3;
         if (jx < 0 || jx >= d.N1 || jy < 0 || jy >= d.N2 || jz < 0 || jz >= d.N3) continue;
         const int jslot = lattice_slot(d, j0, jx + d.N1 * (jy + d.N2 * jz));
         if (dedup) {
            bool exis = false;
            for (int l = 0; l < n; l++) if (nl[(size_t)l * d.Npad + s] == jslot) exis = true;
            if (exis) continue;
         }
         nl[(size_t)n * d.Npad + s] = jslot;
         if (cp)
            for (int a = 0; a < ncomp; a++)
               cp[((size_t)a * z + n) * d.Npad + s] = coupling[(size_t)(i0 * maxslot + q) * ncomp + a];
         n++;
      }
   }
   for (int l = n; l < z; l++) {
      nl[(size_t)l * d.Npad + s] = s;
      if (cp) for (int a = 0; a < ncomp; a++) cp[((size_t)a * z + l) * d.Npad + s] = 0.0;
   }
   count[s] = real ? n : 0;
}

// reduced mode precondition: every atom carries exactly the list length of its ham row
__global__ void lattice_check_kernel(int Npad, const int* __restrict__ ham, const int* __restrict__ count,
                                     const int* __restrict__ lsize, int* __restrict__ bad) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= Npad) return;
   const int h = ham[s];
   if (h >= 0 && count[s] != lsize[h]) atomicAdd(bad, 1);
}

// back-conversion of a device table to the reference's Fortran layout: list(z,N) 1-based original indices
__global__ void table_export_kernel(int N, int Npad, int z, const int* __restrict__ orig, const int* __restrict__ nl,
                                    const int* __restrict__ count, const int* __restrict__ ham,
                                    const int* __restrict__ lsize, int* __restrict__ list) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= Npad) return;
   const int o = orig[s];
   if (o < 0) return;
   const int n = count ? count[s] : lsize[ham[s]];
   for (int j = 0; j < z; j++) list[(size_t)j + (size_t)z * o] = (j < n) ? orig[nl[(size_t)j * Npad + s]] + 1 : 0;
}

// synthetic start for large runs: e_i = normalize(1, a sin(2 pi h), a cos(2 pi h)), h = frac(i*0.6180339887)
__global__ void tilted_moments_kernel(int Npad, int M, int NA, double amp, const int* __restrict__ orig,
                                      const double* __restrict__ mmom_basis, SpinVec* __restrict__ cur,
                                      SpinVec* __restrict__ pred) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (s >= Npad) return;
   const int o = orig[s];
   SpinVec v;
   if (o < 0) { v.x = 0; v.y = 0; v.z = 1; v.m = 0; }
   else {
      const double t = (double)(o + 1) * 0.6180339887;
      const double h = t - floor(t);
      double sn, cs;
      sincospi(2.0 * h, &sn, &cs);
      const double x = 1.0, y = amp * sn, zc = amp * cs;
      const double nrm = sqrt(x * x + y * y + zc * zc);
      v.x = x / nrm; v.y = y / nrm; v.z = zc / nrm; v.m = mmom_basis[o % NA];
   }
   cur[(size_t)k * Npad + s] = v;
   pred[(size_t)k * Npad + s] = v;
}

}  // namespace asd
