// Setup kernel of the staged tile path: turns the slot-major neighbour table nl[z][Npad] into
//   * ulist[tile][ucap] : the unique slots, ordered by (ham row, original atom index), that the 256 atoms of a tile gather from (plus their own slots),
//   * nl16[zq8][Npad]   : the same neighbour table as 16-bit positions in the tile's ulist, 8 per 16-byte word.
// The per-step kernels then stage emomM of ulist in shared memory once per CTA and never gather from global
// memory (asd_device.cuh: llg_stage_kernel<.,.,.,true>).  What the table means is unchanged: entry j of atom i is
// still nlist(j,i) of the reference (source/Hamiltonian/hamiltoniandatatype.f90:30-33), neighbour order kept.
//
// One CTA per tile.  pass 0 only counts the unique slots (so the host can size ucap); pass 1 builds both arrays.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>

namespace asd {

constexpr int TILE = 256;          // threads per CTA of the build kernel = slots per tile of the one-atom-per-thread stage kernel
                                   // (the run kernel uses tiles of ts = 256, 512 or 1024 slots: whole super-bricks)
constexpr int TILE_HASH = 8192;    // open-addressing table (ints) used to find the unique slots
constexpr int TILE_UMAX = 4096;    // beyond this many unique slots per tile the staged path is not used
constexpr size_t TILE_BUILD_SMEM = (size_t)TILE_HASH * 4 + (size_t)TILE_UMAX * (8 + 4);

// `orig` here is the SORT KEY array (Layout::okey): the original atom index for ordinary layouts, the extended
// local index (halo planes included) for a slab.
// Order of a tile's gather list: by (Hamiltonian row, original atom index), i.e. sublattice-major and then the
// reference's own atom order (x fastest).  For a lattice this makes the neighbours that the 32 lanes of a warp
// (one x-run of cells) read for a given shell a CONTIGUOUS run of the list -- including the part that lies in the
// next brick -- so the shared-memory reads of the stage kernels are bank-conflict free.
// Periodic wrap along x: a tile at the x-edge of the supercell reads the cells x = N1-2, N1-1 as the left neighbours
// of x = 0 -- far away in index order, which would split the warp's run.  For lattice layouts (kna, kn1 > 0: key =
// i0 + kna*(ix + kn1*...)) the x index is therefore rotated so that the tile's own cells sit in the middle of the
// rotated range (first cell at koff = (kn1 - BX) / 2): any N1 >= BX + 2 * reach keeps every neighbour run unsplit.
// x-residue split (Monte Carlo block sweep, asd_mc_block.cuh): with split = px > 1 (sna = NA, sn1 = N1 of the lattice) the
// cells of an x-run are listed residue class by residue class (x mod px), so that the cells of ONE colour of a periodic
// colouring with x-period px -- the lanes that are active together in a colour phase -- are consecutive list positions.
struct TileKeyWrap { int kna, kn1, xref, koff, split, sna, sn1; };

// the DM and BQ tables of the layout (z = 0: absent): their neighbours join the tile's gather list and get their own
// 16-bit position tables, so that the stage kernels read them from shared memory as well
struct TileExtra {
   int zdm, zbq;
   const int* __restrict__ dml;
   const int* __restrict__ bql;
   uint4* __restrict__ dm16;   // [ceil(zdm/8)][Npad]
   uint4* __restrict__ bq16;   // [ceil(zbq/8)][Npad]
};

__device__ __forceinline__ unsigned long long tile_key(const int* __restrict__ ham, const int* __restrict__ orig, int slot,
                                                       const TileKeyWrap& kw) {
   int o = orig[slot];
   if (o < 0) return (0x7fffffull << 40) | (unsigned long long)(unsigned)slot;   // padding slots last
   if (kw.kn1 > 0) {
      const int c = o / kw.kna, kx = c % kw.kn1;
      const int rx = (kx - kw.xref + kw.koff + kw.kn1) % kw.kn1;
      o += kw.kna * (rx - kx);
   }
   if (kw.split > 1) {
      const int c = o / kw.sna, kx = c % kw.sn1;
      const int sx = (kx % kw.split) * ((kw.sn1 + kw.split - 1) / kw.split) + kx / kw.split;
      o += kw.sna * (sx - kx);
   }
   return ((unsigned long long)(unsigned)ham[slot] << 40) | (unsigned long long)(unsigned)o;
}

__global__ void __launch_bounds__(TILE)
tile_gather_kernel(int Nown, int Npad, int z, const int* __restrict__ nl, const int* __restrict__ ham, const int* __restrict__ orig,
                   int pass, int ucap, int* __restrict__ ucount, int* __restrict__ ulist, uint4* __restrict__ nl16, int zq8,
                   int kna, int kn1, int koff, int ts, const TileExtra x, int xsplit = 0, int sna = 0, int sn1 = 0,
                   unsigned short* __restrict__ selfpos = nullptr) {
   extern __shared__ unsigned long long tsm64[];
   unsigned long long* keys = tsm64;                       // [TILE_UMAX] (pass 1)
   int* lst = (int*)(tsm64 + TILE_UMAX);                   // [TILE_UMAX] slot of each key (pass 1)
   int* tab = lst + TILE_UMAX;                             // [TILE_HASH]
   __shared__ int nuniq, over, nfill;
   const int tile = blockIdx.x;
   const int s0 = tile * ts + threadIdx.x, send = min((tile + 1) * ts, Nown);
   TileKeyWrap kw{kna, kn1, 0, koff, xsplit, sna, sn1};
   if (kn1 > 0) { const int o0 = orig[tile * ts]; kw.xref = o0 >= 0 ? (o0 / kna) % kn1 : 0; }
   for (int q = threadIdx.x; q < TILE_HASH; q += TILE) tab[q] = -1;
   if (threadIdx.x == 0) { nuniq = 0; over = 0; nfill = 0; }
   __syncthreads();
   auto insert = [&](int key) {
      unsigned hsh = ((unsigned)key * 2654435761u) >> 19;   // 13 bits
      while (true) {
         if (*(volatile int*)&over) return;
         const int old = atomicCAS(&tab[hsh], -1, key);
         if (old == key) return;
         if (old == -1) { if (atomicAdd(&nuniq, 1) + 1 > TILE_UMAX) over = 1; return; }
         hsh = (hsh + 1) & (TILE_HASH - 1);
      }
   };
   for (int s = s0; s < send; s += TILE) {
      insert(s);
      for (int j = 0; j < z; j++) insert(nl[(size_t)j * Npad + s]);
      for (int j = 0; j < x.zdm; j++) insert(x.dml[(size_t)j * Npad + s]);
      for (int j = 0; j < x.zbq; j++) insert(x.bql[(size_t)j * Npad + s]);
   }
   __syncthreads();
   if (pass == 0) {
      if (threadIdx.x == 0) ucount[tile] = over ? INT_MAX : nuniq;
      return;
   }
   if (over) { if (threadIdx.x == 0) ucount[tile] = INT_MAX; return; }
   // compact, pad to a power of two, bitonic sort of (key, slot) pairs by key
   const int cnt = nuniq;
   int np2 = 32;
   while (np2 < cnt) np2 <<= 1;
   for (int q = threadIdx.x; q < TILE_HASH; q += TILE) {
      const int v = tab[q];
      if (v >= 0) { const int at = atomicAdd(&nfill, 1); lst[at] = v; keys[at] = tile_key(ham, orig, v, kw); }
   }
   for (int q = cnt + threadIdx.x; q < np2; q += TILE) { lst[q] = INT_MAX; keys[q] = ~0ull; }
   __syncthreads();
   for (int kk = 2; kk <= np2; kk <<= 1)
      for (int jj = kk >> 1; jj > 0; jj >>= 1) {
         for (int idx = threadIdx.x; idx < np2; idx += TILE) {
            const int ixj = idx ^ jj;
            if (ixj > idx) {
               const unsigned long long a = keys[idx], b = keys[ixj];
               const bool asc = (idx & kk) == 0;
               if ((a > b) == asc) {
                  keys[idx] = b; keys[ixj] = a;
                  const int t0 = lst[idx]; lst[idx] = lst[ixj]; lst[ixj] = t0;
               }
            }
         }
         __syncthreads();
      }
   for (int q = threadIdx.x; q < ucap; q += TILE) ulist[(size_t)tile * ucap + q] = (q < cnt) ? lst[q] : lst[0];
   if (threadIdx.x == 0) ucount[tile] = cnt;
   auto find = [&](int slot) -> unsigned {
      const unsigned long long key = tile_key(ham, orig, slot, kw);
      int lo = 0, hi = cnt - 1;
      while (lo < hi) {
         const int mid = (lo + hi) >> 1;
         if (keys[mid] < key) lo = mid + 1; else hi = mid;
      }
      return (unsigned)lo;
   };
   for (int s = s0; s < send; s += TILE) {
      const unsigned self = find(s);
      if (selfpos) selfpos[s] = (unsigned short)self;
      for (int q = 0; q < zq8; q++) {
         unsigned v[8];
#pragma unroll
         for (int u = 0; u < 8; u++) {
            const int j = 8 * q + u;
            v[u] = (j < z) ? find(nl[(size_t)j * Npad + s]) : self;
         }
         nl16[(size_t)q * Npad + s] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
      }
      for (int kind = 0; kind < 2; kind++) {
         const int zz = kind ? x.zbq : x.zdm;
         const int* __restrict__ tab = kind ? x.bql : x.dml;
         uint4* __restrict__ out = kind ? x.bq16 : x.dm16;
         for (int q = 0; 8 * q < zz; q++) {
            unsigned v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
               const int j = 8 * q + u;
               v[u] = (j < zz) ? find(tab[(size_t)j * Npad + s]) : self;
            }
            out[(size_t)q * Npad + s] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
         }
      }
   }
}

}  // namespace asd
