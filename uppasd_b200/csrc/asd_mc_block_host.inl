// Host side of the Monte Carlo block sweep (asd_mc_block.cuh): tables, tile colouring, launches.  Included by asd_engine.cu.

// Does the block sweep serve this engine?  Device-built lattice (the periodic colouring needs the stencils), reduced
// Hamiltonian, scalar exchange, not a slab.  mc_layout 2 / ASD_MC_BLOCK=1 force it (tests on small lattices), ASD_MC_BLOCK=0
// switches it off; by default it takes over once a tile-colour class fills the GPU.
static bool mc_block_candidate(const asd_engine* e) {
   if (!e->lattice_built || e->slab.on || e->jtensor || !e->sd.reduced || e->sd.t.z <= 0) return false;
   if (e->mc_layout == 0 || e->mc_layout == 1) return false;
   const char* env = std::getenv("ASD_MC_BLOCK");
   if (env && atoi(env) == 0) return false;
   if (e->mc_layout == 2 || (env && atoi(env) == 1)) return true;
   if (std::getenv("ASD_MC_TILES") && atoi(std::getenv("ASD_MC_TILES")) != 0) return false;
   return (long)e->sd.t.Nown * e->M >= 600000L;
}

template <class K>
static void mc_block_launch(K kernel, dim3 g, size_t smem, cudaStream_t st, const Tables& t, const McParams& p, const McBlock& mb,
                            int first, SpinVec* cur) {
   allow_smem(kernel, smem);
   kernel<<<g, 256, smem, st>>>(t, p, mb, first, cur);
}

static int mc_block_prepare(asd_engine* e) {
   McBlockState& B = e->mcb;
   if (B.tried) return 0;
   B.tried = true; B.on = false;
   Layout& L = e->sd;
   Tables& t = L.t;
   int r;
   if (e->lat_ncol == 0 && (r = lattice_colours(e))) return r;
   if (e->lat_ncol > 64) return 0;
   const LatticeDesc& d = e->lat;
   cudaStream_t st = e->stream;
   const long Npad = L.Npad;
   const int Nown = t.Nown > 0 ? t.Nown : L.Npad;
   const int super = d.NA * d.P * d.SY * d.SZ;
   int sms = 148;
   cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
   // tile size: the whole super-brick when its tile-colour classes still fill the GPU, else one 256-slot brick group
   int ts = (super == 1024 && ((long)Nown / 1024) * e->M >= 16L * sms) ? 1024 : 256;
   const char* tenv = std::getenv("ASD_MC_TS");
   if (tenv && (atoi(tenv) == 256 || (atoi(tenv) == 1024 && super == 1024))) ts = atoi(tenv);
   const bool xs = t.zdm > 0 || t.zbq > 0;
   const int ncpl = t.NH * t.z + (xs ? t.NH * t.zdm * 3 + t.NH * t.zbq : 0);
   TileExtra x;
   memset(&x, 0, sizeof x);
   if (xs) { x.zdm = t.zdm; x.dml = t.dml; x.zbq = t.zbq; x.bql = t.bql; }
   const int xsplit = e->lat_period[0] > 1 ? e->lat_period[0] : 0;
   const bool wrap = d.periodic[0] && d.N1 > d.BX;
   const int kna = wrap ? d.NA : 0, kn1 = wrap ? d.N1 : 0, koff = wrap ? (d.N1 - d.BX) / 2 : 0;
   const int* key = L.d_okey.p ? L.d_okey.p : t.orig;
   const int zq8 = (t.z + 7) / 8;
   CU(cudaFuncSetAttribute(tile_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BUILD_SMEM));
   int ntile = 0, ucap = 0;
   for (int attempt = 0; attempt < 2; attempt++) {
      ntile = (Nown + ts - 1) / ts;
      if ((r = B.ucount.alloc(ntile))) return r;
      tile_gather_kernel<<<ntile, TILE, TILE_BUILD_SMEM, st>>>(Nown, (int)Npad, t.z, t.nl, t.ham, key, 0, 0, B.ucount.p, nullptr, nullptr, zq8,
                                                                kna, kn1, koff, ts, x, xsplit, d.NA, d.N1, nullptr);
      e->launches++;
      CU(cudaGetLastError());
      std::vector<int> cnt(ntile);
      CU(cudaMemcpyAsync(cnt.data(), B.ucount.p, (size_t)ntile * sizeof(int), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      const int mx = *std::max_element(cnt.begin(), cnt.end());
      ucap = ((mx + 31) / 32) * 32;
      B.smem = ((size_t)((ncpl + 3) & ~3) + (size_t)4 * ts + (size_t)3 * ucap) * sizeof(double);
      const size_t limit = (ts == 1024) ? (size_t)112 * 1024 : (size_t)56 * 1024;   // 2 / 4 CTAs per SM
      if (mx <= TILE_UMAX && B.smem <= limit) break;
      if (ts == 1024) { ts = 256; continue; }
      if (mx > TILE_UMAX || B.smem > (size_t)220 * 1024) return 0;   // no locality: keep the colour-major kernels
      break;
   }
   if ((r = B.ulist.alloc((size_t)ntile * ucap))) return r;
   if ((r = B.nl16.alloc((size_t)zq8 * Npad))) return r;
   if ((r = B.selfpos.alloc(Npad))) return r;
   if (x.zdm > 0) { if ((r = B.dm16.alloc((size_t)((x.zdm + 7) / 8) * Npad))) return r; x.dm16 = B.dm16.p; }
   if (x.zbq > 0) { if ((r = B.bq16.alloc((size_t)((x.zbq + 7) / 8) * Npad))) return r; x.bq16 = B.bq16.p; }
   tile_gather_kernel<<<ntile, TILE, TILE_BUILD_SMEM, st>>>(Nown, (int)Npad, t.z, t.nl, t.ham, key, 1, ucap, B.ucount.p, B.ulist.p, B.nl16.p, zq8,
                                                             kna, kn1, koff, ts, x, xsplit, d.NA, d.N1, B.selfpos.p);
   e->launches++;
   CU(cudaGetLastError());
   // atoms of every tile in colour order
   const int ncol = e->lat_ncol;
   if ((r = B.corder.alloc((size_t)ntile * ts))) return r;
   if ((r = B.cstart.alloc((size_t)ntile * (ncol + 1)))) return r;
   mc_block_order_kernel<<<ntile, 64, 0, st>>>(Nown, ts, ncol, e->lat_col.p, B.corder.p, B.cstart.p);
   e->launches++;
   CU(cudaGetLastError());
   // tile graph: A ~ B iff an atom of A has a neighbour in B; greedy colouring in tile order
   const int cap = 64;
   DevBuf<int> d_adj, d_nadj;
   if ((r = d_adj.alloc((size_t)ntile * cap))) return r;
   if ((r = d_nadj.alloc(ntile))) return r;
   const size_t bsm = (size_t)((ntile + 31) / 32) * sizeof(unsigned int);
   if (bsm > 200 * 1024) return 0;
   allow_smem(mc_block_adjacency_kernel, bsm);
   mc_block_adjacency_kernel<<<ntile, 256, bsm, st>>>(ntile, ts, ucap, B.ulist.p, B.ucount.p, cap, d_adj.p, d_nadj.p);
   e->launches++;
   CU(cudaGetLastError());
   std::vector<int> adj((size_t)ntile * cap), nadj(ntile);
   CU(cudaMemcpyAsync(adj.data(), d_adj.p, adj.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(nadj.data(), d_nadj.p, nadj.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   std::vector<std::vector<int>> nb(ntile);
   for (int a = 0; a < ntile; a++) {
      if (nadj[a] > cap) return 0;                      // a tile that touches more than 64 others: no locality
      for (int q = 0; q < nadj[a]; q++) { const int b = adj[(size_t)a * cap + q]; nb[a].push_back(b); nb[b].push_back(a); }
   }
   std::vector<int> tcol(ntile, -1), mark;
   int ntc = 0;
   for (int a = 0; a < ntile; a++) {
      mark.assign(ntc + 1, 0);
      for (int b : nb[a]) if (tcol[b] >= 0) mark[tcol[b]] = 1;
      int c = 0;
      while (c < ntc && mark[c]) c++;
      tcol[a] = c;
      if (c == ntc) ntc++;
   }
   B.class_first.assign(ntc, 0); B.class_count.assign(ntc, 0);
   for (int a = 0; a < ntile; a++) B.class_count[tcol[a]]++;
   for (int c = 1; c < ntc; c++) B.class_first[c] = B.class_first[c - 1] + B.class_count[c - 1];
   B.h_tilelist.assign(ntile, 0);
   {
      std::vector<int> fill(B.class_first);
      for (int a = 0; a < ntile; a++) B.h_tilelist[fill[tcol[a]]++] = a;
   }
   if ((r = B.tilelist.upload(B.h_tilelist, st))) return r;
   B.ts = ts; B.ucap = ucap; B.ncol = ncol; B.ntile = ntile;
   B.on = true;
   if (std::getenv("ASD_DEBUG"))
      fprintf(stderr, "[asd] MC block sweep: tiles of %d slots, %d tiles in %d classes, %d atom colours (period %d %d %d), gather list <= %d, %zu B smem\n",
              ts, ntile, ntc, ncol, e->lat_period[0], e->lat_period[1], e->lat_period[2], ucap, B.smem);
   return 0;
}

static McBlock mc_block_params(const asd_engine* e) {
   const McBlockState& B = e->mcb;
   McBlock mb;
   memset(&mb, 0, sizeof mb);
   mb.ts = B.ts; mb.ucap = B.ucap; mb.ncol = B.ncol;
   mb.ulist = B.ulist.p; mb.ucount = B.ucount.p; mb.nl16 = B.nl16.p; mb.dm16 = B.dm16.p; mb.bq16 = B.bq16.p;
   mb.selfpos = B.selfpos.p; mb.corder = B.corder.p; mb.cstart = B.cstart.p; mb.tilelist = B.tilelist.p;
   return mb;
}

static int mc_sweeps_block(asd_engine* e, McParams& p, long nsweeps, long first_sweep) {
   int r = ensure_layout(e, 1);
   if (r) return r;
   McBlockState& B = e->mcb;
   Layout& L = e->sd;
   const McBlock mb = mc_block_params(e);
   const bool xs = L.t.zdm > 0 || L.t.zbq > 0, hb = p.mode == 'H';
   for (long s = 0; s < nsweeps; s++) {
      p.sweep = (unsigned long long)(first_sweep + s);
      for (size_t c = 0; c < B.class_first.size(); c++) {
         const dim3 g((unsigned)B.class_count[c], (unsigned)e->M);
         const int first = B.class_first[c];
#define ASD_MCB(APT, XS, HB) mc_block_launch(mc_block_kernel<APT, XS, HB>, g, B.smem, e->stream, L.t, p, mb, first, e->cur.p)
         if (B.ts == 1024) {
            if (xs) { if (hb) ASD_MCB(4, true, true); else ASD_MCB(4, true, false); }
            else { if (hb) ASD_MCB(4, false, true); else ASD_MCB(4, false, false); }
         } else {
            if (xs) { if (hb) ASD_MCB(1, true, true); else ASD_MCB(1, true, false); }
            else { if (hb) ASD_MCB(1, false, true); else ASD_MCB(1, false, false); }
         }
#undef ASD_MCB
         e->launches++;
      }
   }
   CU(cudaGetLastError());
   return 0;
}

// the sequential visiting order that reproduces the chain of the block sweep: tile colour -> tile -> atom colour -> atom
static int mc_block_visit_order(asd_engine* e, int* order) {
   McBlockState& B = e->mcb;
   Layout& L = e->sd;
   int r = host_orig(e, L);
   if (r) return r;
   std::vector<unsigned short> co((size_t)B.ntile * B.ts);
   std::vector<int> cs((size_t)B.ntile * (B.ncol + 1));
   CU(cudaMemcpy(co.data(), B.corder.p, co.size() * sizeof(unsigned short), cudaMemcpyDeviceToHost));
   CU(cudaMemcpy(cs.data(), B.cstart.p, cs.size() * sizeof(int), cudaMemcpyDeviceToHost));
   size_t n = 0;
   for (int tile : B.h_tilelist) {
      const int nreal = cs[(size_t)tile * (B.ncol + 1) + B.ncol];
      for (int q = 0; q < nreal; q++) {
         const int o = L.orig[(size_t)tile * B.ts + co[(size_t)tile * B.ts + q]];
         if (o >= 0 && n < (size_t)e->N) order[n++] = o + 1;
      }
   }
   if (n != (size_t)e->N) return fail(-5, "internal: block-sweep visiting order covers %zu of %d atoms", n, e->N);
   return 0;
}
