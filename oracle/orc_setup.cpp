// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's table construction.
// Nothing under uppasd_b200/ may include, link or call this file.
//
// Restates, statement by statement, the Fortran setup code that produces the tables read
// by the per-time-step hot path.  Every function cites the reference lines it follows
// (paths relative to the reference root).  Arrays keep the reference's Fortran shapes:
// column-major storage, atom indices 1-based (0 = "no neighbour").
//
//   geometry        source/System/geometry.f90:144-184 (types), :337-488 (coordinates)
//   symmetry ops    source/Hamiltonian/neighbourmap.f90:361-516   (get_symops)
//   shell expansion source/Hamiltonian/neighbourmap.f90:526-640   (get_fullnnlist)
//   neighbour map   source/Hamiltonian/neighbourmap.f90:32-359    (setup_nm)
//   coupling mount  source/Hamiltonian/hamiltonianinit.f90:985-1157 (setup_neighbour_hamiltonian)
//   anisotropy      source/Hamiltonian/hamiltonianinit.f90:857-976  (setup_anisotropies)
//   moments         source/System/magnetizationinit.f90:545-602 (setup_moment), :234-266 (initmag 3)
//
// Parity pinning: see oracle/README.md (reference goldens reproduced in tests/test_oracle_golden.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>

namespace {

const double dbl_tolerance = (double)1e-14f;  // source/Parameters/parameters.f90:12 (default-real literal stored in a double)

// Fortran mod() for integers (truncating, sign of dividend) == C '%'.
inline int fmod_i(int a, int b) { return a % b; }

// 3x3 inverse used twice in the reference (geometry.f90:388-400, neighbourmap.f90:103-119);
// invmatrix(r,c) stored as inv[r-1][c-1].
void inv_cell(const double* C1, const double* C2, const double* C3, double inv[3][3]) {
   double det = C1[0] * C2[1] * C3[2] - C1[0] * C2[2] * C3[1] + C1[1] * C2[2] * C3[0] - C1[1] * C2[0] * C3[2] +
                C1[2] * C2[0] * C3[1] - C1[2] * C2[1] * C3[0];
   for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) inv[a][b] = 0.0;
   if (std::fabs(det) > dbl_tolerance) {
      inv[0][0] = (C2[1] * C3[2] - C3[1] * C2[2]) / det;
      inv[0][1] = (C1[2] * C3[1] - C3[2] * C1[1]) / det;
      inv[0][2] = (C1[1] * C2[2] - C2[1] * C1[2]) / det;
      inv[1][0] = (C2[2] * C3[0] - C3[2] * C2[0]) / det;
      inv[1][1] = (C1[0] * C3[2] - C3[0] * C1[2]) / det;
      inv[1][2] = (C1[2] * C2[0] - C2[2] * C1[0]) / det;
      inv[2][0] = (C2[0] * C3[1] - C3[0] * C2[1]) / det;
      inv[2][1] = (C1[1] * C3[0] - C3[1] * C1[0]) / det;
      inv[2][2] = (C1[0] * C2[1] - C2[0] * C1[1]) / det;
   }
}

struct SymOps {
   int nsym = 0;
   std::vector<std::array<double, 9>> m;  // m[s][ (i-1) + 3*(j-1) ] = sym_mats(i,j,s)
   double& at(int i, int j, int s) { return m[s - 1][(i - 1) + 3 * (j - 1)]; }
};

// neighbourmap.f90:361-516.  Only the built-in point groups (isym 0..3); isym 4/5 read a file.
SymOps get_symops(int isym) {
   SymOps S;
   auto alloc = [&](int n) {
      S.m.assign(n, std::array<double, 9>{});
      for (auto& a : S.m) a.fill(0.0);
   };
   auto sgn = [](int p) { return (p % 2 == 0) ? 1.0 : -1.0; };  // (-1.0)**p
   if (isym == 0) {
      alloc(1);
      for (int i = 1; i <= 3; i++) S.at(i, i, 1) = 1.0;
      S.nsym = 1;
   } else if (isym == 1) {
      alloc(48);
      int c = 0;
      for (int i = 1; i <= 3; i++)
         for (int j = 0; j <= 1; j++) {
            int j_s = (j == 0) ? 1 : -1;
            for (int x = 0; x <= 1; x++)
               for (int y = 0; y <= 1; y++)
                  for (int z = 0; z <= 1; z++) {
                     c++;
                     S.at(1, fmod_i(i - j_s, 3) + 1, c) = sgn(x);
                     S.at(2, fmod_i(i, 3) + 1, c) = sgn(y);
                     S.at(3, fmod_i(i + j_s, 3) + 1, c) = sgn(z);
                  }
         }
      S.nsym = c;
   } else if (isym == 2) {
      alloc(12);
      int c = 0;
      for (int j = 0; j <= 1; j++)
         for (int x = 0; x <= 1; x++)
            for (int y = 0; y <= 1; y++) {
               c++;
               S.at(1, fmod_i(j, 2) + 1, c) = sgn(x);
               S.at(2, fmod_i(j + 1, 2) + 1, c) = sgn(y);
               S.at(3, 3, c) = 1.0;
            }
      S.nsym = c;
   } else if (isym == 3) {
      alloc(24);
      int c = 0;
      double half = 0.5, roothalf = std::sqrt(3.0) * 0.5;
      for (int x = 0; x <= 1; x++)
         for (int y = 0; y <= 1; y++)
            for (int z = 0; z <= 1; z++) {
               c++;
               S.at(1, 1, c) = sgn(x);
               S.at(2, 2, c) = sgn(y);
               S.at(3, 3, c) = sgn(z);
            }
      for (int x1 = 0; x1 <= 1; x1++)
         for (int x2 = 0; x2 <= 1; x2++)
            for (int y1 = 0; y1 <= 1; y1++)
               for (int y2 = 0; y2 <= 1; y2++)
                  if (sgn(x1) * sgn(x2) * sgn(y1) * sgn(y2) < 0.0)
                     for (int z = 0; z <= 1; z++) {
                        c++;
                        S.at(1, 1, c) = sgn(x1) * half;
                        S.at(2, 1, c) = sgn(x2) * roothalf;
                        S.at(1, 2, c) = sgn(y1) * roothalf;
                        S.at(2, 2, c) = sgn(y2) * half;
                        S.at(3, 3, c) = sgn(z);
                     }
      S.nsym = c;
   } else {
      std::fprintf(stderr, "orc: symmetry %d needs sym.mat (not restated)\n", isym);
      std::exit(2);
   }
   return S;
}

struct NeighbourMap {
   int Natom, NT, NA, maxNN, max_no_equiv, max_no_neigh;
   std::vector<int> nm;     // nm(Natom, maxNN, max_no_equiv)
   std::vector<int> nmdim;  // nmdim(maxNN, Natom)
   // unit-cell stencil (kept for inspection by tests)
   std::vector<int> nm_cell;   // (max_no_equiv, max_no_shells, NA)
   std::vector<int> nm_trunk;  // (3, max_no_equiv, max_no_shells, NA)
   std::vector<int> nnm_cell;  // (max_no_shells, NA)
   int max_no_shells;
};

}  // namespace

extern "C" {

// geometry.f90:401-416 (fold basis into the first cell), :437-488 (coordinates, block_size=1),
// :144-184 (atype/anumb).  Bas is updated in place exactly like the reference.
void orc_setup_geometry(int NA, int N1, int N2, int N3, const double* C1, const double* C2, const double* C3,
                        double* Bas /*(3,NA)*/, const int* atype_inp, const int* anumb_inp,
                        double* coord /*(3,Natom)*/, int* atype, int* anumb) {
   double inv[3][3];
   inv_cell(C1, C2, C3, inv);
   for (int i0 = 0; i0 < NA; i0++) {
      double* b = Bas + 3 * i0;
      double icvec[3], bsf[3];
      icvec[0] = b[0] * inv[0][0] + b[1] * inv[1][0] + b[2] * inv[2][0];
      icvec[1] = b[0] * inv[0][1] + b[1] * inv[1][1] + b[2] * inv[2][1];
      icvec[2] = b[0] * inv[0][2] + b[1] * inv[1][2] + b[2] * inv[2][2];
      for (int a = 0; a < 3; a++) bsf[a] = std::floor(icvec[a] + 1e-5);
      for (int a = 0; a < 3; a++) b[a] = b[a] - bsf[0] * C1[a] - bsf[1] * C2[a] - bsf[2] * C3[a];
   }
   long i = 0;
   for (int I3 = 0; I3 < N3; I3++)
      for (int I2 = 0; I2 < N2; I2++)
         for (int I1 = 0; I1 < N1; I1++)
            for (int I0 = 0; I0 < NA; I0++) {
               for (int a = 0; a < 3; a++) coord[3 * i + a] = I1 * C1[a] + I2 * C2[a] + I3 * C3[a] + Bas[3 * I0 + a];
               atype[i] = atype_inp[I0];
               anumb[i] = anumb_inp[I0];
               i++;
            }
}

// neighbourmap.f90:32-359 (setup_nm) with get_symops/get_fullnnlist.  do_ralloy = 0, block_size = 1.
// redcoord is (NT, max_no_shells, 3) column-major, nntype (NT, max_no_shells) or NULL (absent optional).
void* orc_nm_create(int Natom, int NT, int NA, int N1, int N2, int N3, const double* C1, const double* C2,
                    const double* C3, char BC1, char BC2, char BC3, const int* atype, const double* Bas,
                    int max_no_shells, int sym, const int* nn, const double* redcoord, const int* nntype) {
   auto* R = new NeighbourMap();
   R->Natom = Natom; R->NT = NT; R->NA = NA; R->max_no_shells = max_no_shells;
   const double tol = 0.01;
   double inv[3][3];
   inv_cell(C1, C2, C3, inv);
   SymOps S = get_symops(sym);
   int max_no_equiv = S.nsym;  // maxval(nsym); all types share the op set for isym 0..3
   int maxNN = 0;
   for (int t = 0; t < NT; t++) maxNN = std::max(maxNN, nn[t]);
   R->maxNN = maxNN; R->max_no_equiv = max_no_equiv;
   auto RED = [&](int it, int ish, int k) { return redcoord[(it - 1) + NT * ((ish - 1) + (long)max_no_shells * (k - 1))]; };
   // nncoord(3, max_no_equiv, max_no_shells, NT), nmdimt(maxNN, NA) [indexed (ishell,itype) in the reference]
   std::vector<double> nncoord((size_t)3 * max_no_equiv * max_no_shells * NT, 0.0);
   auto NNC = [&](int k, int e, int ish, int it) -> double& {
      return nncoord[(k - 1) + 3 * ((e - 1) + (size_t)max_no_equiv * ((ish - 1) + (size_t)max_no_shells * (it - 1)))];
   };
   std::vector<int> nmdimt((size_t)maxNN * std::max(NA, NT), 0);
   auto NMDIMT = [&](int ish, int it) -> int& { return nmdimt[(ish - 1) + (size_t)maxNN * (it - 1)]; };
   // get_fullnnlist, nelem = 1 (neighbourmap.f90:563-637)
   for (int itype = 1; itype <= NT; itype++)
      for (int ishell = 1; ishell <= nn[itype - 1]; ishell++) {
         if (S.nsym == 1) {
            for (int k = 1; k <= 3; k++) NNC(k, 1, ishell, itype) = RED(itype, ishell, k);
            NMDIMT(ishell, itype) = 1;
         } else {
            int counter = 0;
            for (int isym = 1; isym <= S.nsym; isym++) {
               double tv[3] = {0, 0, 0};
               bool unique = true;
               for (int i = 1; i <= 3; i++)
                  for (int j = 1; j <= 3; j++) tv[i - 1] = tv[i - 1] + RED(itype, ishell, j) * S.at(i, j, isym);
               for (int k = 1; k <= counter; k++) {
                  double d0 = tv[0] - NNC(1, k, ishell, itype), d1 = tv[1] - NNC(2, k, ishell, itype),
                         d2 = tv[2] - NNC(3, k, ishell, itype);
                  if (d0 * d0 + d1 * d1 + d2 * d2 < tol) unique = false;
               }
               if (unique) {
                  counter++;
                  for (int i = 1; i <= 3; i++) NNC(i, counter, ishell, itype) = tv[i - 1];
               }
            }
            NMDIMT(ishell, itype) = counter;
         }
      }
   // first-cell stencil (neighbourmap.f90:161-243)
   R->nm_cell.assign((size_t)max_no_equiv * max_no_shells * NA, 0);
   R->nm_trunk.assign((size_t)3 * max_no_equiv * max_no_shells * NA, 0);
   R->nnm_cell.assign((size_t)max_no_shells * NA, 0);
   auto NMC = [&](int c, int ish, int i0) -> int& { return R->nm_cell[(c - 1) + (size_t)max_no_equiv * ((ish - 1) + (size_t)max_no_shells * (i0 - 1))]; };
   auto NMT = [&](int k, int c, int ish, int i0) -> int& {
      return R->nm_trunk[(k - 1) + 3 * ((c - 1) + (size_t)max_no_equiv * ((ish - 1) + (size_t)max_no_shells * (i0 - 1)))];
   };
   auto NNMC = [&](int ish, int i0) -> int& { return R->nnm_cell[(ish - 1) + (size_t)max_no_shells * (i0 - 1)]; };
   const bool multi = (long)N1 * N2 * N3 > 1;
   for (int I0 = 1; I0 <= NA; I0++) {
      int itype = atype[I0 - 1];
      for (int ishell = 1; ishell <= nn[itype - 1]; ishell++) {
         int counter = 0;
         for (int inei = 1; inei <= NMDIMT(ishell, itype); inei++) {
            double cvec[3], icvec[3], bsf[3], rvec[3];
            for (int a = 0; a < 3; a++) cvec[a] = NNC(a + 1, inei, ishell, itype) + Bas[3 * (I0 - 1) + a];
            icvec[0] = cvec[0] * inv[0][0] + cvec[1] * inv[1][0] + cvec[2] * inv[2][0];
            icvec[1] = cvec[0] * inv[0][1] + cvec[1] * inv[1][1] + cvec[2] * inv[2][1];
            icvec[2] = cvec[0] * inv[0][2] + cvec[1] * inv[1][2] + cvec[2] * inv[2][2];
            for (int a = 0; a < 3; a++) bsf[a] = std::floor(icvec[a] + 5.0e-5);
            for (int a = 0; a < 3; a++) rvec[a] = cvec[a] - bsf[0] * C1[a] - bsf[1] * C2[a] - bsf[2] * C3[a];
            for (int ia = 1; ia <= NA; ia++) {
               int jtype = atype[ia - 1];
               bool typematch = true;
               if (nntype) typematch = (jtype == nntype[(atype[I0 - 1] - 1) + NT * (ishell - 1)]);
               double d0 = rvec[0] - Bas[3 * (ia - 1)], d1 = rvec[1] - Bas[3 * (ia - 1) + 1], d2 = rvec[2] - Bas[3 * (ia - 1) + 2];
               bool tolmatch = d0 * d0 + d1 * d1 + d2 * d2 < tol;
               if (tolmatch && typematch) {
                  counter++;
                  if (max_no_equiv >= counter) {
                     NMC(counter, ishell, I0) = ia;
                     if (multi) {
                        NMT(1, counter, ishell, I0) = (int)std::lround(bsf[0]);
                        NMT(2, counter, ishell, I0) = (int)std::lround(bsf[1]);
                        NMT(3, counter, ishell, I0) = (int)std::lround(bsf[2]);
                     }
                  }
               }
            }
         }
         NNMC(ishell, I0) = counter;
      }
   }
   // full map (neighbourmap.f90:248-321)
   R->nm.assign((size_t)Natom * maxNN * max_no_equiv, 0);
   R->nmdim.assign((size_t)maxNN * Natom, 0);
   auto NM = [&](long iat, int ish, int e) -> int& { return R->nm[(iat - 1) + (size_t)Natom * ((ish - 1) + (size_t)maxNN * (e - 1))]; };
   for (int iz = 0; iz < N3; iz++)
      for (int iy = 0; iy < N2; iy++)
         for (int ix = 0; ix < N1; ix++)
            for (int i0 = 1; i0 <= NA; i0++) {
               int itype = atype[i0 - 1];
               long iat = i0 + (long)ix * NA + (long)iy * N1 * NA + (long)iz * N2 * N1 * NA;
               for (int ishell = 1; ishell <= nn[itype - 1]; ishell++) {
                  R->nmdim[(ishell - 1) + (size_t)maxNN * (iat - 1)] = NMDIMT(ishell, atype[iat - 1]);
                  for (int inei = 1; inei <= NNMC(ishell, i0); inei++) {
                     // NOTE: the reference reads nm_cell/nm_trunk at positions up to nnm_cell, which may exceed
                     // max_no_equiv only if more than max_no_equiv atoms matched; keep the same guard-free read
                     // semantics by clamping to stored entries.
                     if (inei > max_no_equiv) break;
                     int xc_hop = 0, yc_hop = 0, zc_hop = 0;
                     if (multi) { xc_hop = NMT(1, inei, ishell, i0); yc_hop = NMT(2, inei, ishell, i0); zc_hop = NMT(3, inei, ishell, i0); }
                     int j0 = NMC(inei, ishell, i0);
                     int jx = xc_hop + ix;
                     if (BC1 == 'P') jx = fmod_i(jx + 1000 * N1, N1); else if (!multi) jx = 0;
                     int jy = yc_hop + iy;
                     if (BC2 == 'P') jy = fmod_i(jy + 1000 * N2, N2); else if (!multi) jy = 0;
                     int jz = zc_hop + iz;
                     if (BC3 == 'P') jz = fmod_i(jz + 1000 * N3, N3); else if (!multi) jz = 0;
                     if (jx >= 0 && jx < N1 && jy >= 0 && jy < N2 && jz >= 0 && jz < N3) {
                        long jat = j0 + (long)jx * NA + (long)jy * N1 * NA + (long)jz * N2 * N1 * NA;
                        NM(iat, ishell, inei) = (int)jat;
                     }
                  }
               }
            }
   // max_no_neigh (neighbourmap.f90:342-350)
   int mnn = 1;
   for (int i0 = 1; i0 <= NA; i0++) {
      int itype = atype[i0 - 1], counter = 0;
      for (int ishell = 1; ishell <= nn[itype - 1]; ishell++) counter += NNMC(ishell, i0);
      mnn = std::max(mnn, counter);
   }
   R->max_no_neigh = mnn;
   return R;
}

int orc_nm_max_no_neigh(void* h) { return ((NeighbourMap*)h)->max_no_neigh; }
int orc_nm_max_no_equiv(void* h) { return ((NeighbourMap*)h)->max_no_equiv; }
int orc_nm_maxnn(void* h) { return ((NeighbourMap*)h)->maxNN; }
void orc_nm_get(void* h, int* nm, int* nmdim) {
   auto* R = (NeighbourMap*)h;
   std::memcpy(nm, R->nm.data(), R->nm.size() * sizeof(int));
   std::memcpy(nmdim, R->nmdim.data(), R->nmdim.size() * sizeof(int));
}
// unit-cell stencil: nm_cell (max_no_equiv,max_no_shells,NA), nm_trunk (3,max_no_equiv,max_no_shells,NA), nnm_cell (max_no_shells,NA)
void orc_nm_get_stencil(void* h, int* nm_cell, int* nm_trunk, int* nnm_cell) {
   auto* R = (NeighbourMap*)h;
   std::memcpy(nm_cell, R->nm_cell.data(), R->nm_cell.size() * sizeof(int));
   std::memcpy(nm_trunk, R->nm_trunk.data(), R->nm_trunk.size() * sizeof(int));
   std::memcpy(nnm_cell, R->nnm_cell.data(), R->nnm_cell.size() * sizeof(int));
}
void orc_nm_free(void* h) { delete (NeighbourMap*)h; }

// hamiltonianinit.f90:985-1157 (setup_neighbour_hamiltonian), do_ralloy = 0, conf_num = 1, no LSF.
//   xc(hdim, NT, max_no_shells); ammom_inp(NA); nlist(max_no_neigh, Natom) must arrive zeroed
//   (hamiltoniandata.f90:111); ncoup(hdim, max_no_neigh, nHam); nlistsize(nHam).
void orc_mount(void* h, int Natom, int NT, int NA, int nHam, const int* anumb, const int* atype, int max_no_neigh,
               const int* nn, const double* xc, const double* ammom_inp, int hdim, int lexp, int do_sortcoup,
               int map_multiple, double mry, double mub, int* nlistsize, int* nlist, double* ncoup) {
   auto* R = (NeighbourMap*)h;
   const int max_no_shells = R->max_no_shells, maxNN = R->maxNN;
   auto NM = [&](long iat, int ish, int e) { return R->nm[(iat - 1) + (size_t)Natom * ((ish - 1) + (size_t)maxNN * (e - 1))]; };
   auto NMDIM = [&](int ish, long iat) { return R->nmdim[(ish - 1) + (size_t)maxNN * (iat - 1)]; };
   auto NL = [&](int l, long i) -> int& { return nlist[(l - 1) + (size_t)max_no_neigh * (i - 1)]; };
   auto NC = [&](int a, int l, long i) -> double& { return ncoup[(a - 1) + (size_t)hdim * ((l - 1) + (size_t)max_no_neigh * (i - 1))]; };
   auto XC = [&](int a, int it, int k) { return xc[(a - 1) + (size_t)hdim * ((it - 1) + (size_t)NT * (k - 1))]; };
   for (size_t q = 0; q < (size_t)hdim * max_no_neigh * nHam; q++) ncoup[q] = 0.0;
   const double fc2 = 2.0 * mry / mub;
   for (long i = 1; i <= Natom; i++) {
      int ncount = 1;
      for (int k = 1; k <= nn[atype[i - 1] - 1]; k++)
         for (int j = 1; j <= NMDIM(k, i); j++) {
            int nb = NM(i, k, j);
            if (nb > 0) {
               bool exis = false;
               for (int l = 1; l <= ncount - 1; l++)
                  if (NL(l, i) == nb) exis = true;
               if (!exis || map_multiple) {
                  NL(ncount, i) = nb;
                  if (i <= nHam) {
                     double mi = ammom_inp[anumb[i - 1] - 1], mj = ammom_inp[anumb[nb - 1] - 1];
                     // abs(...) < 1e-6 : default-real literal compared against a double
                     if (std::fabs(mi * mj) < (double)1e-6f) {
                        for (int a = 1; a <= hdim; a++) NC(a, ncount, i) = 0.0;
                     } else {
                        // xc * fc2 / m_i**lexp / m_j**lexp, left to right
                        double pi_ = (lexp == 2) ? mi * mi : mi, pj_ = (lexp == 2) ? mj * mj : mj;
                        for (int a = 1; a <= hdim; a++) NC(a, ncount, i) = XC(a, atype[i - 1], k) * fc2 / pi_ / pj_;
                     }
                  }
                  ncount++;
               }
            }
         }
      if (i <= nHam) nlistsize[i - 1] = ncount - 1;
   }
   (void)max_no_shells; (void)NA;
   if (do_sortcoup && nHam == Natom) {
      std::vector<double> tmp(hdim);
      for (long i = 1; i <= Natom; i++)
         for (int j = 1; j <= nlistsize[i - 1]; j++)
            for (int k = 1; k <= nlistsize[i - 1] - j; k++)
               if (NL(k, i) > NL(k + 1, i)) {
                  int t = NL(k, i); NL(k, i) = NL(k + 1, i); NL(k + 1, i) = t;
                  for (int a = 1; a <= hdim; a++) { double c = NC(a, k, i); NC(a, k, i) = NC(a, k + 1, i); NC(a, k + 1, i) = c; }
               }
   }
}

// hamiltonianinit.f90:985-1157 (setup_neighbour_hamiltonian), the do_ralloy = 1 branch (:1075-1084) for a NON-DILUTE random alloy
// (every site occupied: Natom = Natom_full, acellnumb is the identity, so setup_nm's map is the one of do_ralloy = 0,
// neighbourmap.f90:257-262,309).  conf_num = 1, no LSF, nHam = Natom.
//   xc(hdim, NT, max_no_shells, Nchmax, Nchmax); ammom_inp(NA, Nchmax); atype_ch / asite_ch / achem_ch (Natom):
//   ncoup(:, ncount, i) = xc(:, atype_ch(i), k, achem_ch(i), achem_ch(j)) * fc2 / m(asite_ch(i), achem_ch(i))**lexp
//                                                                                / m(asite_ch(j), achem_ch(j))**lexp
void orc_mount_alloy(void* h, int Natom, int NT, int NA, int Nchmax, const int* atype_ch, const int* asite_ch, const int* achem_ch,
                     int max_no_neigh, const int* nn, const double* xc, const double* ammom_inp, int hdim, int lexp,
                     int do_sortcoup, int map_multiple, double mry, double mub, int* nlistsize, int* nlist, double* ncoup) {
   auto* R = (NeighbourMap*)h;
   const int max_no_shells = R->max_no_shells, maxNN = R->maxNN;
   auto NM = [&](long iat, int ish, int e) { return R->nm[(iat - 1) + (size_t)Natom * ((ish - 1) + (size_t)maxNN * (e - 1))]; };
   auto NMDIM = [&](int ish, long iat) { return R->nmdim[(ish - 1) + (size_t)maxNN * (iat - 1)]; };
   auto NL = [&](int l, long i) -> int& { return nlist[(l - 1) + (size_t)max_no_neigh * (i - 1)]; };
   auto NC = [&](int a, int l, long i) -> double& { return ncoup[(a - 1) + (size_t)hdim * ((l - 1) + (size_t)max_no_neigh * (i - 1))]; };
   auto XC = [&](int a, int it, int k, int ci, int cj) {
      return xc[(a - 1) + (size_t)hdim * ((it - 1) + (size_t)NT * ((k - 1) + (size_t)max_no_shells * ((ci - 1) + (size_t)Nchmax * (cj - 1))))];
   };
   auto AM = [&](int site, int chem) { return ammom_inp[(site - 1) + (size_t)NA * (chem - 1)]; };
   for (size_t q = 0; q < (size_t)hdim * max_no_neigh * Natom; q++) ncoup[q] = 0.0;
   const double fc2 = 2.0 * mry / mub;
   for (long i = 1; i <= Natom; i++) {
      int ncount = 1;
      for (int k = 1; k <= nn[atype_ch[i - 1] - 1]; k++)
         for (int j = 1; j <= NMDIM(k, i); j++) {
            int nb = NM(i, k, j);
            if (nb > 0) {
               bool exis = false;
               for (int l = 1; l <= ncount - 1; l++)
                  if (NL(l, i) == nb) exis = true;
               if (!exis || map_multiple) {
                  NL(ncount, i) = nb;
                  const double mi = AM(asite_ch[i - 1], achem_ch[i - 1]), mj = AM(asite_ch[nb - 1], achem_ch[nb - 1]);
                  if (std::fabs(mi * mj) < (double)1e-6f) {
                     for (int a = 1; a <= hdim; a++) NC(a, ncount, i) = 0.0;
                  } else {
                     double pi_ = (lexp == 2) ? mi * mi : mi, pj_ = (lexp == 2) ? mj * mj : mj;
                     for (int a = 1; a <= hdim; a++)
                        NC(a, ncount, i) = XC(a, atype_ch[i - 1], k, achem_ch[i - 1], achem_ch[nb - 1]) * fc2 / pi_ / pj_;
                  }
                  ncount++;
               }
            }
         }
      nlistsize[i - 1] = ncount - 1;
   }
   if (do_sortcoup) {
      for (long i = 1; i <= Natom; i++)
         for (int j = 1; j <= nlistsize[i - 1]; j++)
            for (int k = 1; k <= nlistsize[i - 1] - j; k++)
               if (NL(k, i) > NL(k + 1, i)) {
                  int t = NL(k, i); NL(k, i) = NL(k + 1, i); NL(k + 1, i) = t;
                  for (int a = 1; a <= hdim; a++) { double c = NC(a, k, i); NC(a, k, i) = NC(a, k + 1, i); NC(a, k + 1, i) = c; }
               }
   }
}

// hamiltonianinit.f90:895-913 (setup_anisotropies, do_ralloy=0, no random anisotropy, mult_axis='N').
// anisotropytype(NA), anisotropy(NA,6) column-major; outputs taniso(N), eaniso(3,N), kaniso(2,N), sb(N).
void orc_setup_anisotropies(int Natom, int NA, const int* anumb, const int* anisotropytype, const double* anisotropy,
                            const double* ammom_inp, double mry, double mub, int* taniso, double* eaniso,
                            double* kaniso, double* sb) {
   const double fc = mry / mub;
   auto AN = [&](int ia, int c) { return anisotropy[(ia - 1) + (size_t)NA * (c - 1)]; };
   for (long i = 1; i <= Natom; i++) {
      int ia = anumb[i - 1];
      taniso[i - 1] = anisotropytype[ia - 1];
      double anorm = AN(ia, 3) * AN(ia, 3) + AN(ia, 4) * AN(ia, 4) + AN(ia, 5) * AN(ia, 5);
      double den = std::sqrt(anorm + 1.0e-15);
      for (int a = 0; a < 3; a++) eaniso[3 * (i - 1) + a] = AN(ia, 3 + a) / den;
      double m = ammom_inp[ia - 1];
      // integer powers expanded the way gfortran's powi does: m**4=(m*m)**2, m**6=((m*m)*m)**2
      double m2 = m * m, m4 = m2 * m2, m3 = m2 * m, m6 = m3 * m3;
      if (taniso[i - 1] == 2) {
         kaniso[2 * (i - 1)] = fc * AN(ia, 1) / m4;
         kaniso[2 * (i - 1) + 1] = fc * AN(ia, 2) / m6;
      } else {
         kaniso[2 * (i - 1)] = fc * AN(ia, 1) / m2;
         kaniso[2 * (i - 1) + 1] = fc * AN(ia, 2) / m4;
      }
      sb[i - 1] = AN(ia, 6);
   }
}

// magnetizationinit.f90:545-602 (setup_moment) and :234-266 (initmag 3), do_ralloy = 0.
void orc_setup_moments(int Natom, int Mensemble, int NA, int N1, int N2, int N3, const double* ammom_inp,
                       const double* aemom_inp /*(3,NA) normalised*/, const double* Landeg_ch, double* mmom,
                       double* mmom0, double* mmomi, double* Landeg, double* emom, double* emomM) {
   for (int I3 = 0; I3 < N3; I3++)
      for (int I2 = 0; I2 < N2; I2++)
         for (int I1 = 0; I1 < N1; I1++)
            for (int I0 = 1; I0 <= NA; I0++) {
               long i = I0 + (long)I1 * NA + (long)I2 * N1 * NA + (long)I3 * N2 * N1 * NA;
               double m = std::fabs(ammom_inp[I0 - 1]);
               Landeg[i - 1] = Landeg_ch[I0 - 1] * 0.5;
               for (int j = 0; j < Mensemble; j++) {
                  size_t q = (i - 1) + (size_t)Natom * j;
                  mmom[q] = m; mmom0[q] = m; mmomi[q] = 1.0 / m;
               }
            }
   for (int I3 = 0; I3 < N3; I3++)
      for (int I2 = 0; I2 < N2; I2++)
         for (int I1 = 0; I1 < N1; I1++)
            for (int I0 = 1; I0 <= NA; I0++) {
               long i = I0 + (long)I1 * NA + (long)I2 * N1 * NA + (long)I3 * N2 * N1 * NA;
               for (int j = 0; j < Mensemble; j++) {
                  // emomM = aemom_inp * mmom(anumb(I0), j): anumb(I0) == I0 in the first cell
                  double mm = mmom[(I0 - 1) + (size_t)Natom * j];
                  for (int a = 0; a < 3; a++) {
                     emom[a + 3 * ((i - 1) + (size_t)Natom * j)] = aemom_inp[a + 3 * (I0 - 1)];
                     emomM[a + 3 * ((i - 1) + (size_t)Natom * j)] = aemom_inp[a + 3 * (I0 - 1)] * mm;
                  }
               }
            }
}

}  // extern "C"
