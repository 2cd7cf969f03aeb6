// Host side of the Monte Carlo block sweep (asd_mc_block.cuh): tables, tile colouring, launches.  Included by asd_engine.cu.

// Does the block sweep serve this engine?  Device-built lattice (the periodic colouring needs the stencils), reduced
// Hamiltonian, scalar exchange, not a slab.  mc_layout 2 / ASD_MC_BLOCK=1 force it (tests on small lattices), ASD_MC_BLOCK=0
// switches it off; by default it takes over once a tile-colour class fills the GPU AND every tile is in the run form (see the
// end of mc_block_prepare).
static bool mc_block_candidate(const asd_engine* e) {
   if (!e->lattice_built || e->slab.on || e->jtensor || !e->sd.reduced || e->sd.t.z <= 0) return false;
   if (e->mc_layout == 0 || e->mc_layout == 1) return false;
   const char* env = std::getenv("ASD_MC_BLOCK");
   if (env && atoi(env) == 0) return false;
   if (e->mc_layout == 2 || (env && atoi(env) == 1)) return true;
   if (std::getenv("ASD_MC_TILES") && atoi(std::getenv("ASD_MC_TILES")) != 0) return false;
   return (long)e->sd.t.Nown * e->M >= 400000L;
}

template <class K>
static void mc_block_run_launch(K kernel, dim3 g, size_t smem, cudaStream_t st, const Tables& t, const McParams& p, const McBlock& mb,
                                const McRuns& mr, const McTicket& tk, int first, SpinVec* cur) {
   allow_smem(kernel, smem);
   kernel<<<g, 256, smem, st>>>(t, p, mb, mr, tk, first, cur);
}

template <class K>
static void mc_block_ws_launch(K kernel, dim3 g, size_t smem, cudaStream_t st, const Tables& t, const McParams& p, const McBlock& mb,
                               const McRuns& mr, const McTicket& tk, SpinVec* cur) {
   allow_smem(kernel, smem);
   kernel<<<g, 512, smem, st>>>(t, p, mb, mr, tk, cur);
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
   // tile size: the whole super-brick (the run form with whole-sweep scheduling needs it: every tile of the sweep is in one launch, so
   // two CTAs per SM worth of tiles is enough to fill the GPU), else one 256-slot brick group
   int ts = (super == 1024 && ((long)Nown / 1024) * e->M >= 2L * sms) ? 1024 : 256;
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
   // RUN form (asd_mc_runs.cuh): neighbour arrangement per Hamiltonian row, group table + per-tile verification; regular tiles
   // come first in every class
   B.class_run.assign(ntc, 0);
   B.smem_run = 0;
   {
      const char* renv = std::getenv("ASD_MC_RUNS");
      const char* nenv = std::getenv("ASD_MC_NT");
      B.nt = (nenv && atoi(nenv) == 512) ? 512 : 256;
      if (!xs && ts == 1024 && ncol <= 56 && t.z <= MCR_ZMAX && !(renv && atoi(renv) == 0)) {
         McRuns& mr = B.mr;
         memset(&mr, 0, sizeof mr);
         mr.q = 2;
         mr.batch = (B.nt == 512) ? MCR_BATCH_WS : MCR_BATCH;
         const int spcap = MCR_ZMAX;
         DevBuf<unsigned short> d_arr;
         DevBuf<double> d_cseq;
         DevBuf<int> d_steps, d_ok;
         if ((r = d_arr.alloc((size_t)t.NH * mr.q * spcap))) return r;
         if ((r = d_cseq.alloc((size_t)t.NH * spcap))) return r;
         if ((r = d_steps.alloc(t.NH))) return r;
         mc_runs_arrange_kernel<<<(t.NH + 31) / 32, 32, 0, st>>>(t.NH, t.z, mr.q, spcap, t.lsize, t.cp, d_arr.p, d_cseq.p, d_steps.p);
         e->launches++;
         CU(cudaGetLastError());
         std::vector<int> steps(t.NH);
         std::vector<double> cseq((size_t)t.NH * spcap);
         CU(cudaMemcpyAsync(steps.data(), d_steps.p, steps.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
         CU(cudaMemcpyAsync(cseq.data(), d_cseq.p, cseq.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
         CU(cudaStreamSynchronize(st));
         const int smin = *std::min_element(steps.begin(), steps.end()), smax = *std::max_element(steps.begin(), steps.end());
         mr.sp = (std::max(smax, 1) + 3) & ~3;
         // worth it only while the padding of the arrangement stays small: steps x lanes <= list length + 25 %
         const bool fits = smin >= 0 && (long)t.NH * mr.sp <= MCR_CSEQ && (long)smax * mr.q <= (long)t.z + t.z / 4 + mr.q;
         if (fits) {
            for (int h = 0; h < t.NH; h++)
               for (int q = 0; q < mr.sp; q++) mr.cseq[h * mr.sp + q] = cseq[(size_t)h * spcap + q];
            mr.ncw = (MCR_PH0 + 4 * MCR_PHMAX + 7) & ~7;
            mr.gstride = (mr.ncw + MCR_GMAX * (MCR_HDR + mr.q * mr.sp) + 7) & ~7;
            mr.zero = ucap;
            if ((r = B.gtab.alloc((size_t)ntile * mr.gstride))) return r;
            if ((r = d_ok.alloc(ntile))) return r;
            mc_block_groups_kernel<<<ntile, 256, 0, st>>>(ts, ncol, t.z, (size_t)Npad, t.ham, B.nl16.p, B.selfpos.p, B.corder.p, B.cstart.p,
                                                           B.ucount.p, d_arr.p, d_steps.p, spcap, mr, B.gtab.p, d_ok.p);
            e->launches++;
            CU(cudaGetLastError());
            std::vector<int> ok(ntile);
            CU(cudaMemcpyAsync(ok.data(), d_ok.p, (size_t)ntile * sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            const int gmax = *std::max_element(ok.begin(), ok.end());
            mr.gwords = (mr.ncw + gmax * (MCR_HDR + mr.q * mr.sp) + 7) & ~7;
            mr.gtab = B.gtab.p;
            // 256 threads: one record buffer of MCR_BATCH attempts; 512 threads (warp-specialised): two of MCR_BATCH_WS -- the same bytes
            B.smem_run = (size_t)5 * MCR_BATCH * sizeof(double) + (size_t)mr.gwords * 2 + (size_t)MCR_BATCH * 2 +
                         (size_t)3 * (ucap + 16) * sizeof(double);
            if (gmax > 0 && B.smem_run <= (size_t)113 * 1024) {
               for (int c = 0; c < ntc; c++) {
                  int* lst = B.h_tilelist.data() + B.class_first[c];
                  B.class_run[c] = (int)(std::stable_partition(lst, lst + B.class_count[c], [&](int a) { return ok[a] > 0; }) - lst);
               }
            }
         }
      }
   }
   if ((r = B.tilelist.upload(B.h_tilelist, st))) return r;
   // whole-sweep scheduling (McTicket): every tile in the run form, symmetric tile adjacency of at most 64 neighbours
   B.ticket = false;
   {
      int nrun = 0, deg = 0;
      for (int c : B.class_run) nrun += c;
      for (int a = 0; a < ntile; a++) {
         std::sort(nb[a].begin(), nb[a].end());
         nb[a].erase(std::unique(nb[a].begin(), nb[a].end()), nb[a].end());
         deg = std::max(deg, (int)nb[a].size());
      }
      const char* tenv2 = std::getenv("ASD_MC_TICKET");
      if (nrun == ntile && deg <= 64 && ntc <= 255 && !(tenv2 && atoi(tenv2) == 0)) {
         B.adjcap = std::max(deg, 1);
         std::vector<int> hadj((size_t)ntile * B.adjcap, 0), hn(ntile);
         std::vector<unsigned char> hc(ntile);
         for (int a = 0; a < ntile; a++) {
            hn[a] = (int)nb[a].size();
            hc[a] = (unsigned char)tcol[a];
            std::copy(nb[a].begin(), nb[a].end(), hadj.begin() + (size_t)a * B.adjcap);
         }
         if ((r = B.adj.upload(hadj, st))) return r;
         if ((r = B.nadj.upload(hn, st))) return r;
         if ((r = B.tclass.upload(hc, st))) return r;
         if ((r = B.done.alloc((size_t)ntile * e->M))) return r;
         if ((r = B.counter.alloc(1))) return r;
         CU(cudaMemsetAsync(B.done.p, 0, (size_t)ntile * e->M * sizeof(unsigned int), st));
         CU(cudaMemsetAsync(B.counter.p, 0, sizeof(unsigned long long), st));
         B.tickets = 0; B.epoch = 0;
         B.ticket = true;
      }
   }
   B.ts = ts; B.ucap = ucap; B.ncol = ncol; B.ntile = ntile;
   // By default the block sweep takes over only in its run form with whole-sweep scheduling (measured: bcc Fe 128^3 9.2e9 attempts/s
   // against 6.4e9 of the colour-major launches).  The generic per-atom-word form (DM / BQ tables, irregular tiles) is slower than the
   // colour-major launches on the layouts measured (2-D triangular + DMI 1024^2 x 2: 6.8e9 vs 1.2e10) and runs on request only.
   {
      const char* env = std::getenv("ASD_MC_BLOCK");
      const bool forced = e->mc_layout == 2 || (env && atoi(env) == 1);
      B.on = forced || B.ticket;
   }
   if (std::getenv("ASD_DEBUG"))
   {
      int nrun = 0;
      for (int c : B.class_run) nrun += c;
      fprintf(stderr, "[asd] MC block sweep: tiles of %d slots, %d tiles in %d classes, %d atom colours (period %d %d %d), gather list <= %d, %zu B smem; "
                      "run form on %d tiles (%d threads, %d steps x %d lanes, %zu B smem)%s\n",
              ts, ntile, ntc, ncol, e->lat_period[0], e->lat_period[1], e->lat_period[2], ucap, B.smem, nrun, B.nt, B.mr.sp, B.mr.q, B.smem_run, B.ticket ? ", one launch per sweep (tickets)" : "");
   }
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
   McTicket tk;
   memset(&tk, 0, sizeof tk);
   // ASD_MC_PREDRAW=1: trial moves of a sweep drawn for the whole lattice by one data-parallel launch (mc_predraw_kernel) before the
   // sweep kernel(s) of the run form (256-thread kernel on 1024-slot tiles).  NOT the default: measured SLOWER on bcc 128^3 (0.532
   // against 0.442 ms per sweep, profiles/r03p): inside the sweep CTAs the draws overlap the other CTA's colour phases and fill the
   // dependency wait, while the separate launch adds 168 MB of record traffic each way and its own divergent transcendental code
   bool any_run = B.ticket;
   for (int nr : B.class_run) any_run = any_run || nr > 0;
   const char* penv = std::getenv("ASD_MC_PREDRAW");
   const bool predraw = (penv && atoi(penv) != 0) && any_run && B.nt == 256 && B.ts == 1024;
   McRuns mr = B.mr;
   mr.drec = nullptr; mr.ntile = B.ntile;
   if (predraw) {
      if ((r = B.drec.alloc((size_t)e->M * B.ntile * 1024 * 5))) return r;
      mr.drec = B.drec.p;
   }
   for (long s = 0; s < nsweeps; s++) {
      p.sweep = (unsigned long long)(first_sweep + s);
      if (predraw) {
         const dim3 gp((unsigned)(((size_t)B.ntile * 1024 + 255) / 256), (unsigned)e->M);
         if (hb) mc_predraw_kernel<true><<<gp, 256, 0, e->stream>>>(L.t, p, B.corder.p, B.ntile, e->cur.p, B.drec.p);
         else mc_predraw_kernel<false><<<gp, 256, 0, e->stream>>>(L.t, p, B.corder.p, B.ntile, e->cur.p, B.drec.p);
         e->launches++;
      }
      if (B.ticket) {
         // one launch for the whole sweep: tickets in class order, dependencies through done[] (asd_mc_runs.cuh)
         tk.counter = B.counter.p; tk.base = B.tickets; tk.epoch = ++B.epoch; tk.done = B.done.p; tk.adj = B.adj.p; tk.nadj = B.nadj.p;
         tk.tclass = B.tclass.p; tk.cap = B.adjcap; tk.ntile = B.ntile;
         {
            const char* la = std::getenv("ASD_MC_LOOKAHEAD");
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
            tk.lookahead = la ? atoi(la) : sms;       // measured on bcc 128^3: 0 -> 0.456, 148 -> 0.443, 296 -> 0.451, 592 -> 0.469 ms per sweep
         }
         const dim3 gr((unsigned)((long)B.ntile * e->M));
         if (B.nt == 512) {
            if (hb) mc_block_ws_launch(mc_block_ws_kernel<true>, gr, B.smem_run, e->stream, L.t, p, mb, B.mr, tk, e->cur.p);
            else mc_block_ws_launch(mc_block_ws_kernel<false>, gr, B.smem_run, e->stream, L.t, p, mb, B.mr, tk, e->cur.p);
         } else {
            if (hb) mc_block_run_launch(mc_block_run_kernel<true, true>, gr, B.smem_run, e->stream, L.t, p, mb, mr, tk, 0, e->cur.p);
            else mc_block_run_launch(mc_block_run_kernel<false, true>, gr, B.smem_run, e->stream, L.t, p, mb, mr, tk, 0, e->cur.p);
         }
         B.tickets += (unsigned long long)B.ntile * e->M;
         e->launches++;
         continue;
      }
      for (size_t c = 0; c < B.class_first.size(); c++) {
         const int nrun = B.class_run[c];
         if (nrun > 0) {
            const dim3 gr((unsigned)nrun, (unsigned)e->M);
            if (hb) mc_block_run_launch(mc_block_run_kernel<true, false>, gr, B.smem_run, e->stream, L.t, p, mb, mr, tk, B.class_first[c], e->cur.p);
            else mc_block_run_launch(mc_block_run_kernel<false, false>, gr, B.smem_run, e->stream, L.t, p, mb, mr, tk, B.class_first[c], e->cur.p);
            e->launches++;
         }
         if (nrun == B.class_count[c]) continue;
         const dim3 g((unsigned)(B.class_count[c] - nrun), (unsigned)e->M);
         const int first = B.class_first[c] + nrun;
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
