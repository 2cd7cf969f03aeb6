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
   int reduced;      // 1: couplings per basis atom (ham row = basis atom), 0: per atom
   int Ncell;        // N1*N2*N3
   int N, Npad;
   // Device order = BRICK order: the supercell is cut into bricks of BX x BY x BZ cells (P = BX*BY*BZ, a multiple
   // of 32); a brick holds NA runs of P slots, one per basis atom (so a warp always works on one sublattice),
   // cells x-fastest inside the brick.  One 256-thread CTA then covers a compact 3-D block of atoms whose
   // neighbour sets overlap heavily: the set of distinct spins a tile gathers from stays small enough to be
   // staged in shared memory (asd_tiles.cuh) instead of being gathered from L2 fifty times.
   int BX, BY, BZ, P;
   int NTX, NTY, NTZ;   // bricks per direction (the last one may be partly empty -> padding slots)
   // SUPER-BRICKS: SY x SZ bricks that are adjacent in y / z are numbered consecutively, so that SY*SZ consecutive
   // 256-slot tiles form one compact block (e.g. 32 x 4 x 4 cells): the big tiles of the run kernel (asd_runs.cuh)
   // stage ~40 % fewer spins per atom than four separate bricks.  NSY, NSZ = super-bricks per direction.
   int SY, SZ, NSY, NSZ;
   // Slab decomposition along z, one slab per GPU: this engine owns the global planes
   // [z0, z0 + N3) of N3g; N1, N2, N3, Ncell, N describe the LOCAL slab.  Slots [0, Nown) are the owned bricks;
   // the H halo planes of the lower ring neighbour and then of the upper one follow.  Halo slots hold copies of
   // the neighbours' boundary spins: gathered from, never computed (orig = -1).
   int slab, z0, N3g, H, Nown;
   int has_lo, has_hi;  // a neighbour exists below / above (always, when z is periodic)
};

// halo slot of basis atom i0 in cell (ix,iy) of halo plane hz (0..H-1) on side 0 (below) / 1 (above)
__device__ __host__ __forceinline__ int halo_slot(const LatticeDesc& d, int side, int hz, int i0, int ix, int iy) {
   return d.Nown + ((((side * d.H + hz) * d.NA + i0) * d.N2 + iy) * d.N1 + ix);
}

// slot of basis atom i0 in cell (ix,iy,iz)
__device__ __host__ __forceinline__ int lattice_slot(const LatticeDesc& d, int i0, int ix, int iy, int iz) {
   const int tx = ix / d.BX, ty = iy / d.BY, tz = iz / d.BZ;
   const int lx = ix - tx * d.BX, ly = iy - ty * d.BY, lz = iz - tz * d.BZ;
   const int sbk = tx + d.NTX * (ty / d.SY + d.NSY * (tz / d.SZ));
   const int brick = sbk * (d.SY * d.SZ) + (ty % d.SY) + d.SY * (tz % d.SZ);
   return (brick * d.NA + i0) * d.P + lx + d.BX * (ly + d.BY * lz);
}

// inverse: slot -> (i0, ix, iy, iz); returns false for a padding slot (cell outside the supercell)
__device__ __host__ __forceinline__ bool lattice_unslot(const LatticeDesc& d, int s, int& i0, int& ix, int& iy, int& iz) {
   const int run = s / d.P, c = s - run * d.P;
   const int brick = run / d.NA;
   i0 = run - brick * d.NA;
   const int sbk = brick / (d.SY * d.SZ), q = brick - sbk * (d.SY * d.SZ);
   const int tx = sbk % d.NTX, ty = ((sbk / d.NTX) % d.NSY) * d.SY + q % d.SY, tz = (sbk / (d.NTX * d.NSY)) * d.SZ + q / d.SY;
   const int lx = c % d.BX, ly = (c / d.BX) % d.BY, lz = c / (d.BX * d.BY);
   ix = tx * d.BX + lx; iy = ty * d.BY + ly; iz = tz * d.BZ + lz;
   return ix < d.N1 && iy < d.N2 && iz < d.N3;
}

// orig[] / ham[] of every device slot (original atom index = i0 + NA*(ix + N1*(iy + N2*iz)), geometry.f90:337-488),
// the sort key of the tile gather lists (okey: atom index extended over the halo planes, so that x-runs of a
// slab's halo stay contiguous in the lists) and, for a slab, the halo slots of the neighbours that mirror an atom.
__global__ void lattice_index_kernel(const LatticeDesc d, int* __restrict__ orig, int* __restrict__ ham, int* __restrict__ okey,
                                     int* __restrict__ hdst_lo, int* __restrict__ hdst_hi) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= d.Npad) return;
   int i0, ix, iy, iz, o = -1, h = -1, key = -1;
   if (s < d.Nown) {
      if (lattice_unslot(d, s, i0, ix, iy, iz)) {
         o = i0 + d.NA * (ix + d.N1 * (iy + d.N2 * iz));
         h = d.reduced ? i0 : 0;
         key = i0 + d.NA * (ix + d.N1 * (iy + d.N2 * (iz + d.H)));
         if (hdst_lo) {
            // planes [0,H) are the upper halo (side 1) of the lower neighbour, planes [N3-H,N3) the lower halo of the upper one
            hdst_lo[s] = (d.slab && d.has_lo && iz < d.H) ? halo_slot(d, 1, iz, i0, ix, iy) : -1;
            hdst_hi[s] = (d.slab && d.has_hi && iz >= d.N3 - d.H) ? halo_slot(d, 0, iz - (d.N3 - d.H), i0, ix, iy) : -1;
         }
      } else if (hdst_lo) { hdst_lo[s] = -1; hdst_hi[s] = -1; }
   } else {
      int q = s - d.Nown;
      if (q < 2 * d.H * d.NA * d.N2 * d.N1) {
         ix = q % d.N1; q /= d.N1;
         iy = q % d.N2; q /= d.N2;
         i0 = q % d.NA; q /= d.NA;
         const int hz = q % d.H, side = q / d.H;
         const int lz = side == 0 ? hz - d.H : d.N3 + hz;   // local plane index, outside [0, N3)
         h = d.reduced ? i0 : 0;
         key = i0 + d.NA * (ix + d.N1 * (iy + d.N2 * (lz + d.H)));
      }
   }
   orig[s] = o;
   ham[s] = h;
   okey[s] = key;
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
   int i0 = 0, ix = 0, iy = 0, iz = 0;
   const bool real = s < d.Nown && lattice_unslot(d, s, i0, ix, iy, iz);
   int n = 0;
   if (real) {
      const int ns = nslot[i0];
      for (int q = 0; q < ns; q++) {
         const int j0 = cell_atom[i0 * maxslot + q] - 1;
         const int* sh = cell_shift + 3 * (i0 * maxslot + q);
         int jx = sh[0] + ix, jy = sh[1] + iy, jz = sh[2] + iz;
         if (d.periodic[0]) jx = (jx + 1000 * d.N1) % d.N1;
         if (d.periodic[1]) jy = (jy + 1000 * d.N2) % d.N2;
         if (jx < 0 || jx >= d.N1 || jy < 0 || jy >= d.N2) continue;
         int jslot;
         if (!d.slab) {
            if (d.periodic[2]) jz = (jz + 1000 * d.N3) % d.N3;
            if (jz < 0 || jz >= d.N3) continue;
            jslot = lattice_slot(d, j0, jx, jy, jz);
         } else {
            // slab: jz is the UNWRAPPED local plane; the neighbour exists iff its global plane does
            const int gz = d.z0 + jz;
            if (!d.periodic[2] && (gz < 0 || gz >= d.N3g)) continue;
            if (jz >= 0 && jz < d.N3) jslot = lattice_slot(d, j0, jx, jy, jz);
            else if (jz < 0 && jz >= -d.H) jslot = halo_slot(d, 0, jz + d.H, j0, jx, jy);
            else if (jz >= d.N3 && jz < d.N3 + d.H) jslot = halo_slot(d, 1, jz - d.N3, j0, jx, jy);
            else continue;   // cannot happen: the host sizes H from the stencil
         }
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
__global__ void tilted_moments_kernel(int Nown, int Npad, int M, int NA, double amp, unsigned int atom_offset, const int* __restrict__ orig,
                                      const double* __restrict__ mmom_basis, SpinVec* __restrict__ cur,
                                      SpinVec* __restrict__ pred) {
   const int s = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
   if (s >= Nown) return;   // halo slots of a slab are filled by the neighbours
   const int o = orig[s];
   SpinVec v;
   if (o < 0) { v.x = 0; v.y = 0; v.z = 1; v.m = 0; }
   else {
      const double t = (double)((unsigned int)o + atom_offset + 1u) * 0.6180339887;
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
