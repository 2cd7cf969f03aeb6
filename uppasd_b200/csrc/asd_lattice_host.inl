// C ABI: on-device table construction + table export (included by asd_engine.cu)

static int host_orig(asd_engine* e, Layout& L) {
   if (!L.orig.empty()) return 0;
   L.orig.resize(L.Npad);
   CU(cudaMemcpy(L.orig.data(), L.d_orig.p, (size_t)L.Npad * sizeof(int), cudaMemcpyDeviceToHost));
   L.slot_of.assign(e->N, -1);
   for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) L.slot_of[L.orig[s]] = s;
   return 0;
}

// Brick shape of the device order (LatticeDesc): P = BX*BY*BZ cells, a multiple of 32, with NA*P close to one
// 256-thread tile.  Cost model (shared-memory wavefronts per tile of the staged stage kernel, measured with ncu on
// bcc Fe, profiles/): the gather loop costs 2 wavefronts per 8-byte LDS when a warp is one 32-cell x-run (its
// neighbours are then at most two contiguous runs of the gather list -> conflict-free), ~3 for 16-cell and ~3.6 for
// 8-cell rows; staging costs ~0.44 wavefronts per distinct spin, which grows with the brick's halo.
static void choose_brick(LatticeDesc& d) {
   const int Nd[3] = {d.N1, d.N2, d.N3};
   int target = 256 / d.NA;
   int p2 = 32;
   while (p2 * 2 <= target) p2 *= 2;
   target = p2;
   // reduced Hamiltonians with 4 or 8 basis atoms: 128 cells per brick, so that every 128-slot group (4 x-runs, the
   // unit of the run kernel) is one sublattice and 2 bricks / 1 brick make a 1024-slot tile
   if (d.reduced && (d.NA == 4 || d.NA == 8)) target = 128;
   double best = 1e300;
   int bb[3] = {32, 1, 1};
   for (int bx = 1; bx <= 256; bx *= 2)
      for (int by = 1; by <= 256; by *= 2)
         for (int bz = 1; bz <= 256; bz *= 2) {
            const int P = bx * by * bz;
            if (P % 32 != 0 || P > target) continue;
            const int b[3] = {bx, by, bz};
            double pad = 1.0, halo = 1.0;
            for (int a = 0; a < 3; a++) {
               pad *= (double)((Nd[a] + b[a] - 1) / b[a]) * b[a] / Nd[a];
               halo *= (double)(b[a] + (Nd[a] > 1 ? 4 : 0)) / b[a];
            }
            const double wf = (bx >= 32) ? 2.0 : (bx == 16) ? 3.0 : (bx == 8) ? 3.6 : 4.0;
            const double loop = wf * 150.0;             // 50 neighbours x 3 loads, per atom
            const double stage = 0.44 * 0.55 * halo;    // distinct spins per atom ~ 0.55 x bounding box
            const double score = pad * (loop + stage) * (1.0 + 0.02 * std::log2((double)target / P));
            if (score < best) { best = score; bb[0] = bx; bb[1] = by; bb[2] = bz; }
         }
   d.BX = bb[0]; d.BY = bb[1]; d.BZ = bb[2]; d.P = bb[0] * bb[1] * bb[2];
   d.NTX = (d.N1 + d.BX - 1) / d.BX; d.NTY = (d.N2 + d.BY - 1) / d.BY; d.NTZ = (d.N3 + d.BZ - 1) / d.BZ;
   // super-bricks (big tiles of the run kernel): 1024 slots = 4, 2 or 1 bricks, when the Hamiltonian is reduced (the run
   // kernel's precondition) and every 128-slot group is one sublattice.  ASD_SUPER = "sy,sz" overrides (1,1 = off).
   d.SY = d.SZ = 1;
   if (d.reduced && (d.NA * d.P == 256 || (d.P % 128 == 0 && 1024 % (d.NA * d.P) == 0))) {
      const int nb = 1024 / (d.NA * d.P);
      if (nb == 4) {
         if (d.NTZ >= 2 && d.NTY >= 2) { d.SY = 2; d.SZ = 2; }
         else if (d.NTY >= 4) { d.SY = 4; d.SZ = 1; }
         else if (d.NTZ >= 4) { d.SY = 1; d.SZ = 4; }
      } else if (nb == 2) {
         if (d.NTY >= 2) d.SY = 2; else if (d.NTZ >= 2) d.SZ = 2;
      }
      const char* env = std::getenv("ASD_SUPER");
      int sy, sz;
      if (env && sscanf(env, "%d,%d", &sy, &sz) == 2 && sy >= 1 && sz >= 1 && sy * sz <= nb) { d.SY = sy; d.SZ = sz; }
   }
   d.NSY = (d.NTY + d.SY - 1) / d.SY; d.NSZ = (d.NTZ + d.SZ - 1) / d.SZ;
}

// shape, brick order and slot counts of a supercell (slab settings from e->slab)
static int fill_lattice_desc(asd_engine* e, int NA, int N1, int N2, int N3l, int N3g, const int* periodic, bool reduced) {
   LatticeDesc& d = e->lat;
   const Slab& sb = e->slab;
   memset(&d, 0, sizeof d);
   d.NA = NA; d.N1 = N1; d.N2 = N2; d.N3 = N3l;
   for (int a = 0; a < 3; a++) d.periodic[a] = periodic[a];
   d.reduced = reduced ? 1 : 0;
   d.Ncell = N1 * N2 * N3l;
   d.N = e->N;
   d.slab = sb.on; d.N3g = N3g; d.z0 = sb.on ? sb.g * N3l : 0; d.H = sb.on ? sb.H : 0;
   d.has_lo = sb.on && (d.periodic[2] || sb.g > 0);
   d.has_hi = sb.on && (d.periodic[2] || sb.g < sb.G - 1);
   choose_brick(d);
   const long nown = (long)d.NTX * d.NSY * d.NSZ * d.SY * d.SZ * d.NA * d.P;
   const long np = ((nown + 2L * d.H * NA * N1 * N2 + 31) / 32) * 32;
   if (np > 2000000000L) return fail(-3, "too many atoms for 32-bit device indices");
   d.Nown = (int)nown; d.Npad = (int)np;
   return 0;
}

extern "C" {

int asd_build_lattice_table(asd_engine* e, int kind, int NA, int N1, int N2, int N3, const char* bc3, int maxslot,
                            const int* nslot, const int* cell_atom, const int* cell_shift, const double* coupling) {
   if (e->N == 0) return fail(-2, "asd_set_system must be called first");
   CU(cudaSetDevice(e->device));
   if (!(e->NH == NA || e->NH == e->N)) return fail(-1, "nHam must be NA (do_reduced Y) or Natom");
   if (kind < 0 || kind > 3) return fail(-1, "kind must be 0 (exchange), 1 (DM), 2 (BQ) or 3 (tensorial exchange)");
   // kind 3: the exchange table with nine couplings per pair (do_jtensor 1, hamiltonianinit.f90:412-432): same map, same
   // buffers as kind 0, J(a,b) at component a + 3 b
   const bool tensor = (kind == 3);
   if (tensor) kind = 0;
   const int ncomp = tensor ? 9 : (kind == 1) ? 3 : 1;
   Layout& L = e->sd;
   LatticeDesc& d = e->lat;
   Slab& sb = e->slab;
   // slab decomposition: N3 is the GLOBAL plane count, this engine owns N3 / nslabs planes starting at z0
   const int N3l = sb.on ? N3 / sb.G : N3;
   if (sb.on && (N3 % sb.G != 0 || N3l < sb.H)) return fail(-1, "slab: N3 = %d must be a multiple of the %d slabs and each slab at least %d planes thick", N3, sb.G, sb.H);
   if ((long)NA * N1 * N2 * N3l != e->N) return fail(-1, "NA*N1*N2*N3%s = %ld does not match Natom = %d", sb.on ? "/nslabs" : "", (long)NA * N1 * N2 * N3l, e->N);
   e->tri_layout = 0;
   if (!e->lattice_built) {
      int per[3];
      for (int a = 0; a < 3; a++) per[a] = (bc3[a] == 'P' || bc3[a] == 'p') ? 1 : 0;
      {
         int rr = fill_lattice_desc(e, NA, N1, N2, N3l, N3, per, e->NH < e->N);
         if (rr) return rr;
      }
      if (d.reduced)
         for (int i = 0; i < e->N; i++) if (e->aHam[i] != i % NA + 1) return fail(-1, "aHam is not the basis-atom number; cannot use the lattice builder");
      L.N = e->N; L.Npad = d.Npad; L.NH = e->NH; L.M = e->M; L.reduced = d.reduced;
      L.orig.clear(); L.slot_of.clear();
      int r;
      if ((r = L.d_orig.alloc(d.Npad))) return r;
      if ((r = L.d_ham.alloc(d.Npad))) return r;
      if ((r = L.d_okey.alloc(d.Npad))) return r;
      if (sb.on) { if ((r = sb.hdst_lo.alloc(d.Nown))) return r; if ((r = sb.hdst_hi.alloc(d.Nown))) return r; }
      lattice_index_kernel<<<(d.Npad + 255) / 256, 256, 0, e->stream>>>(d, L.d_orig.p, L.d_ham.p, L.d_okey.p, sb.hdst_lo.p, sb.hdst_hi.p);
      e->launches++;
      memset(&L.t, 0, sizeof L.t);
      L.t.N = e->N; L.t.Npad = d.Npad; L.t.Nown = d.Nown; L.t.M = e->M; L.t.NH = e->NH; L.t.reduced = d.reduced;
      L.t.atom_offset = (unsigned int)((long)NA * N1 * N2 * d.z0);
      L.t.ham = L.d_ham.p; L.t.orig = L.d_orig.p;
      L.t.ext_uniform = 1;
      e->lattice_built = true;
      e->ex = HostTable(); e->dm = HostTable(); e->bq = HostTable();
   } else if (d.NA != NA || d.N1 != N1 || d.N2 != N2 || d.N3 != N3l) return fail(-1, "lattice differs from the first asd_build_lattice_table call");
   if (sb.on)
      for (int i0 = 0; i0 < NA; i0++)
         for (int q = 0; q < nslot[i0] && q < maxslot; q++)
            if (std::abs(cell_shift[3 * (i0 * maxslot + q) + 2]) > sb.H)
               return fail(-1, "slab: the stencil reaches %d planes along z but the halo holds %d", std::abs(cell_shift[3 * (i0 * maxslot + q) + 2]), sb.H);
   int z = 1;
   for (int i0 = 0; i0 < NA; i0++) { if (nslot[i0] < 0 || nslot[i0] > maxslot) return fail(-1, "nslot out of range"); z = std::max(z, nslot[i0]); }
   // can two stencil entries of one basis atom land on the same atom? (small periodic cells) -> need de-duplication
   int dedup = 0;
   const int Nd[3] = {N1, N2, N3};   // GLOBAL extents: the table is the one of the whole supercell
   for (int i0 = 0; i0 < NA && !dedup; i0++)
      for (int a = 0; a < nslot[i0] && !dedup; a++)
         for (int b = a + 1; b < nslot[i0]; b++) {
            if (cell_atom[i0 * maxslot + a] != cell_atom[i0 * maxslot + b]) continue;
            bool same = true;
            for (int c = 0; c < 3; c++) {
               const int df = cell_shift[3 * (i0 * maxslot + a) + c] - cell_shift[3 * (i0 * maxslot + b) + c];
               if (d.periodic[c] ? (df % Nd[c] != 0) : (df != 0)) same = false;
            }
            if (same) { dedup = 1; break; }
         }
   if (dedup && sb.on) return fail(-1, "slab: supercell too small along some direction (a neighbour appears twice through the periodic wrap)");
   {
      Stencil& S = e->stencil[kind];
      S.maxslot = maxslot;
      S.nslot.assign(nslot, nslot + NA);
      S.cell_atom.assign(cell_atom, cell_atom + (size_t)NA * maxslot);
      S.cell_shift.assign(cell_shift, cell_shift + (size_t)3 * NA * maxslot);
      e->lat_ncol = 0;
      e->mcb.tried = false; e->mcb.on = false;
   }
   DevBuf<int> d_nslot, d_catom, d_cshift;
   DevBuf<double> d_coupl;
   int r;
   if ((r = d_nslot.upload(std::vector<int>(nslot, nslot + NA), e->stream))) return r;
   if ((r = d_catom.upload(std::vector<int>(cell_atom, cell_atom + (size_t)NA * maxslot), e->stream))) return r;
   if ((r = d_cshift.upload(std::vector<int>(cell_shift, cell_shift + (size_t)3 * NA * maxslot), e->stream))) return r;
   if ((r = d_coupl.upload(std::vector<double>(coupling, coupling + (size_t)ncomp * NA * maxslot), e->stream))) return r;
   DevBuf<int>& d_list = (kind == 0) ? L.d_nl : (kind == 1) ? L.d_dml : L.d_bql;
   DevBuf<double>& d_cp = (kind == 0) ? L.d_cp : (kind == 1) ? L.d_dmv : L.d_jbq;
   DevBuf<int>& d_size = (kind == 0) ? L.d_lsize : (kind == 1) ? L.d_dmsize : L.d_bqsize;
   DevBuf<int>& d_cnt = (kind == 0) ? L.d_cnt[0] : (kind == 1) ? L.d_cnt[1] : L.d_cnt[2];
   if ((r = d_list.alloc((size_t)z * d.Npad))) return r;
   if ((r = d_cnt.alloc(d.Npad))) return r;
   if (!d.reduced && (r = d_cp.alloc((size_t)ncomp * z * d.Npad))) return r;
   lattice_table_kernel<<<(d.Npad + 127) / 128, 128, 0, e->stream>>>(d, maxslot, z, ncomp, dedup, d_nslot.p, d_catom.p, d_cshift.p,
                                                                   d_coupl.p, d_list.p, d_cnt.p, d.reduced ? nullptr : d_cp.p);
   e->launches++;
   CU(cudaGetLastError());
   if (d.reduced) {
      // coupling rows and list lengths come from the atoms of the first cell (i <= nHam, hamiltonianinit.f90:1059)
      std::vector<int> lsize(NA, 0);
      std::vector<double> rows((size_t)NA * z * ncomp, 0.0);
      for (int i0 = 0; i0 < NA; i0++) {
         std::vector<long> acc;
         for (int q = 0; q < nslot[i0]; q++) {
            const int* sh = cell_shift + 3 * (i0 * maxslot + q);
            int j[3] = {sh[0], sh[1], sh[2]};
            bool ok = true;
            for (int c = 0; c < 3; c++) {
               if (d.periodic[c]) j[c] = (j[c] + 1000 * Nd[c]) % Nd[c];
               if (j[c] < 0 || j[c] >= Nd[c]) ok = false;
            }
            if (!ok) continue;
            const long jat = (cell_atom[i0 * maxslot + q] - 1) + (long)NA * (j[0] + (long)N1 * (j[1] + (long)N2 * j[2]));
            if (dedup && std::find(acc.begin(), acc.end(), jat) != acc.end()) continue;
            for (int a = 0; a < ncomp; a++) rows[((size_t)i0 * z + acc.size()) * ncomp + a] = coupling[(size_t)(i0 * maxslot + q) * ncomp + a];
            acc.push_back(jat);
         }
         lsize[i0] = (int)acc.size();
      }
      if ((r = d_cp.upload(rows, e->stream))) return r;
      if ((r = d_size.upload(lsize, e->stream))) return r;
      DevBuf<int> bad;
      if ((r = bad.upload(std::vector<int>(1, 0), e->stream))) return r;
      lattice_check_kernel<<<(d.Nown + 255) / 256, 256, 0, e->stream>>>(d.Nown, L.d_ham.p, d_cnt.p, d_size.p, bad.p);
      e->launches++;
      int nbad = 0;
      CU(cudaMemcpyAsync(&nbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      if (nbad) return fail(-4, "do_reduced: %d atoms do not have the complete neighbour set of their basis atom "
                                "(open boundary in a direction that has neighbours)", nbad);
   }
   CU(cudaStreamSynchronize(e->stream));
   Tables& t = L.t;
   if (kind == 0) { t.z = z; t.nl = d_list.p; t.cp = d_cp.p; t.lsize = d_size.p; t.jtens = tensor ? 1 : 0; e->jtensor = tensor; }
   if (kind == 1) { t.zdm = z; t.dml = d_list.p; t.dmv = d_cp.p; t.dmsize = d_size.p; }
   if (kind == 2) { t.zbq = z; t.bql = d_list.p; t.jbq = d_cp.p; t.bqsize = d_size.p; }
   L.zs[kind] = z;
   e->committed = false;
   return 0;
}

int asd_get_table_dims(asd_engine* e, int kind, int* z, int* ncomp) {
   if (kind < 0 || kind > 2) return fail(-1, "bad kind");
   const HostTable& T = (kind == 0) ? e->ex : (kind == 1) ? e->dm : e->bq;
   int zz = e->lattice_built ? e->sd.zs[kind] : T.z;
   if (z) *z = zz;
   if (ncomp) *ncomp = (kind == 1) ? 3 : (kind == 0 && e->jtensor) ? 9 : 1;
   return 0;
}

int asd_get_table(asd_engine* e, int kind, int* list, int* listsize, double* coup) {
   if (kind < 0 || kind > 2) return fail(-1, "bad kind");
   CU(cudaSetDevice(e->device));
   const int ncomp = (kind == 1) ? 3 : (kind == 0 && e->jtensor) ? 9 : 1;
   if (!e->lattice_built) {
      const HostTable& T = (kind == 0) ? e->ex : (kind == 1) ? e->dm : e->bq;
      if (!T.present()) return fail(-2, "table not set");
      if (list) std::copy(T.list.begin(), T.list.end(), list);
      if (listsize) std::copy(T.lsize.begin(), T.lsize.end(), listsize);
      if (coup) std::copy(T.coup.begin(), T.coup.end(), coup);
      return 0;
   }
   Layout& L = e->sd;
   const int z = L.zs[kind];
   if (z == 0) return fail(-2, "table not built");
   DevBuf<int>& d_list = (kind == 0) ? L.d_nl : (kind == 1) ? L.d_dml : L.d_bql;
   DevBuf<double>& d_cp = (kind == 0) ? L.d_cp : (kind == 1) ? L.d_dmv : L.d_jbq;
   DevBuf<int>& d_size = (kind == 0) ? L.d_lsize : (kind == 1) ? L.d_dmsize : L.d_bqsize;
   DevBuf<int>& d_cnt = L.d_cnt[kind];
   int r;
   if (list) {
      DevBuf<int> out;
      if ((r = out.alloc((size_t)z * e->N))) return r;
      CU(cudaMemsetAsync(out.p, 0, (size_t)z * e->N * sizeof(int), e->stream));
      table_export_kernel<<<(L.Npad + 255) / 256, 256, 0, e->stream>>>(e->N, L.Npad, z, L.d_orig.p, d_list.p, d_cnt.p, L.d_ham.p, d_size.p, out.p);
      e->launches++;
      CU(cudaMemcpyAsync(list, out.p, (size_t)z * e->N * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
   }
   if (L.reduced) {
      if (listsize) CU(cudaMemcpy(listsize, d_size.p, (size_t)e->NH * sizeof(int), cudaMemcpyDeviceToHost));
      if (coup) {
         std::vector<double> rows((size_t)e->NH * z * ncomp);
         CU(cudaMemcpy(rows.data(), d_cp.p, rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
         for (int h = 0; h < e->NH; h++)
            for (int j = 0; j < z; j++)
               for (int a = 0; a < ncomp; a++) coup[a + (size_t)ncomp * (j + (size_t)z * h)] = rows[((size_t)h * z + j) * ncomp + a];
      }
   } else {
      if ((r = host_orig(e, L))) return r;
      if (listsize) {
         std::vector<int> cnt(L.Npad);
         CU(cudaMemcpy(cnt.data(), d_cnt.p, (size_t)L.Npad * sizeof(int), cudaMemcpyDeviceToHost));
         for (int s = 0; s < L.Npad; s++) if (L.orig[s] >= 0) listsize[L.orig[s]] = cnt[s];
      }
      if (coup) {
         std::vector<double> c((size_t)ncomp * z * L.Npad);
         CU(cudaMemcpy(c.data(), d_cp.p, c.size() * sizeof(double), cudaMemcpyDeviceToHost));
         for (int s = 0; s < L.Npad; s++) {
            const int o = L.orig[s];
            if (o < 0) continue;
            for (int j = 0; j < z; j++)
               for (int a = 0; a < ncomp; a++) coup[a + (size_t)ncomp * (j + (size_t)z * o)] = c[((size_t)a * z + j) * L.Npad + s];
         }
      }
   }
   return 0;
}

int asd_init_moments_tilted(asd_engine* e, double amplitude, int NA, const double* mmom_basis) {
   if (!e->committed) return fail(-2, "asd_commit has not been called");
   CU(cudaSetDevice(e->device));
   Layout& L = e->sd;
   int r;
   if ((r = e->cur.alloc((size_t)L.Npad * e->M))) return r;
   if ((r = e->pred.alloc((size_t)L.Npad * e->M))) return r;
   DevBuf<double> mb;
   if ((r = mb.upload(std::vector<double>(mmom_basis, mmom_basis + NA), e->stream))) return r;
   dim3 g, b;
   launch_cfg(L.Npad, e->M, g, b);
   tilted_moments_kernel<<<g, b, 0, e->stream>>>(L.t.Nown, L.Npad, e->M, NA, amplitude, L.t.atom_offset, L.d_orig.p, mb.p, e->cur.p, e->pred.p);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(e->stream));
   e->state_layout = 1;
   e->msum_fresh = false;
   e->mm_valid = false;
   if ((r = slab_push_state(e))) return r;
   e->h_emom.clear(); e->h_mmom.clear();
   if (e->mompar != 0) return fail(-1, "mompar != 0 needs asd_set_moments (mmom0)");
   return 0;
}

}  // extern "C"

// device-built tables -> HostTable (reference shapes), for the host-side construction of the Monte Carlo layout
static int materialise_host_tables(asd_engine* e) {
   HostTable* T[3] = {&e->ex, &e->dm, &e->bq};
   for (int kind = 0; kind < 3; kind++) {
      const int z = e->sd.zs[kind];
      if (z == 0 || T[kind]->present()) continue;
      const int nc = (kind == 1) ? 3 : 1;
      HostTable& H = *T[kind];
      H.z = z; H.ncomp = nc;
      H.list.assign((size_t)z * e->N, 0);
      H.lsize.assign(e->NH, 0);
      H.coup.assign((size_t)nc * z * e->NH, 0.0);
      int r = asd_get_table(e, kind, H.list.data(), H.lsize.data(), H.coup.data());
      if (r) { H = HostTable(); return r; }
   }
   return 0;
}

// Periodic colouring of a device-built lattice.  The colour of an atom depends only on its basis atom and on its
// GLOBAL cell coordinates modulo a small period (p1, p2, p3), so every slab of a decomposed supercell derives the
// same colouring without communication.  The quotient graph (p1*p2*p3*NA sites, edges = the union of the stencils of
// all tables, symmetrised) is coloured greedily on the host; a proper colouring of the quotient lifts to a proper
// colouring of the supercell as long as no edge becomes a self-loop of the quotient (checked for every candidate period;
// p_a > max |shift_a| guarantees it).  Periods must divide the extent of periodic directions.  Among the admissible
// periods the one with the fewest colours wins.
static int lattice_colours(asd_engine* e) {
   const LatticeDesc& d = e->lat;
   const int NA = d.NA;
   const int Ng[3] = {d.N1, d.N2, d.slab ? d.N3g : d.N3};
   int reach[3] = {0, 0, 0};
   for (const Stencil& S : e->stencil) {
      if (!S.present()) continue;
      for (int i0 = 0; i0 < NA; i0++)
         for (int q = 0; q < S.nslot[i0]; q++)
            for (int a = 0; a < 3; a++) reach[a] = std::max(reach[a], std::abs(S.cell_shift[3 * (i0 * S.maxslot + q) + a]));
   }
   std::vector<int> cand[3];
   for (int a = 0; a < 3; a++) {
      if (Ng[a] == 1 || reach[a] == 0) { cand[a].push_back(1); continue; }
      // periods not larger than the reach are admissible too when no shift aliases an atom with itself in the quotient (the
      // check below rejects those that do): bcc Fe with 4 shells has reach 2 and the 8-colouring of period (2, 2, 2).  Ties
      // keep the smallest x period (loops ascend): the block sweep lists an x-run residue class by residue class.
      for (int p = 2; p <= std::min(Ng[a], 2 * reach[a] + 4); p++)
         if (!d.periodic[a] || Ng[a] % p == 0) cand[a].push_back(p);
      if (cand[a].empty()) {
         // no small divisor: the whole extent is always a valid period (small supercells may alias neighbours, which
         // only merges edges of the quotient graph)
         if (Ng[a] > 4096) return fail(-5, "no colouring period found along direction %d (extent %d, stencil reach %d)", a + 1, Ng[a], reach[a]);
         cand[a].push_back(Ng[a]);
      }
   }
   int best_ncol = 1 << 30, best_p[3] = {1, 1, 1};
   std::vector<unsigned char> best_col;
   for (int p1 : cand[0]) for (int p2 : cand[1]) for (int p3 : cand[2]) {
      const int P[3] = {p1, p2, p3};
      const int ns = p1 * p2 * p3 * NA;
      auto site = [&](int i0, int c1, int c2, int c3) { return ((c3 * p2 + c2) * p1 + c1) * NA + i0; };
      std::vector<std::vector<int>> adj(ns);
      bool ok = true;
      for (int c3 = 0; c3 < p3 && ok; c3++) for (int c2 = 0; c2 < p2 && ok; c2++) for (int c1 = 0; c1 < p1 && ok; c1++)
         for (int i0 = 0; i0 < NA && ok; i0++) {
            const int u = site(i0, c1, c2, c3);
            for (const Stencil& S : e->stencil) {
               if (!S.present()) continue;
               for (int q = 0; q < S.nslot[i0]; q++) {
                  const int* sh = S.cell_shift.data() + 3 * (i0 * S.maxslot + q);
                  const int j0 = S.cell_atom[i0 * S.maxslot + q] - 1;
                  const int c[3] = {c1, c2, c3};
                  // an open direction shorter than the shift: no cell has this neighbour (e.g. BC 0 along a one-cell axis)
                  bool exists = true;
                  for (int a = 0; a < 3; a++) if (!d.periodic[a] && std::abs(sh[a]) >= Ng[a]) exists = false;
                  if (!exists) continue;
                  int n[3];
                  for (int a = 0; a < 3; a++) n[a] = ((c[a] + sh[a]) % P[a] + P[a]) % P[a];
                  const int v = site(j0, n[0], n[1], n[2]);
                  if (v == u) {
                     // a true self-neighbour (periodic image of the atom itself in a tiny supercell) is harmless;
                     // an alias of a different atom is not
                     bool self = true;
                     for (int a = 0; a < 3; a++) if (sh[a] != 0 && !(d.periodic[a] && sh[a] % Ng[a] == 0)) self = false;
                     if (!self) ok = false;
                     continue;
                  }
                  adj[u].push_back(v);
                  adj[v].push_back(u);
               }
            }
         }
      if (!ok) continue;
      std::vector<int> colour(ns, -1), mark;
      int ncol = 0;
      for (int u = 0; u < ns; u++) {
         mark.assign(ncol + 1, 0);
         for (int v : adj[u]) if (colour[v] >= 0) mark[colour[v]] = 1;
         int c = 0;
         while (c < ncol && mark[c]) c++;
         colour[u] = c;
         if (c == ncol) ncol++;
      }
      if (ncol < best_ncol && ncol < 255) {
         best_ncol = ncol; best_p[0] = p1; best_p[1] = p2; best_p[2] = p3;
         best_col.assign(colour.begin(), colour.end());
      }
   }
   if (best_col.empty()) return fail(-5, "no periodic colouring of the lattice found");
   DevBuf<unsigned char> cellcol;
   int r;
   if ((r = cellcol.upload(best_col, e->stream))) return r;
   if ((r = e->lat_col.alloc(d.Nown))) return r;
   lattice_colour_kernel<<<(d.Nown + 255) / 256, 256, 0, e->stream>>>(d, best_p[0], best_p[1], best_p[2], cellcol.p, e->lat_col.p);
   e->launches++;
   CU(cudaGetLastError());
   CU(cudaStreamSynchronize(e->stream));
   e->lat_ncol = best_ncol;
   for (int a = 0; a < 3; a++) e->lat_period[a] = best_p[a];
   return 0;
}
