// Batched, variable-size multi-site evaluation (dgp_batch_* of include/dgp.h): the B200 replacement of the reference's
// one-worker-per-site map (examples/nwqn-loadest-example/nwqn-loadest-example.py:156-157) for sites that share a model.
//
// Every launch of the single-site schedule (dgp_api.cu) covers ALL sites of the batch: the tile engine
// maps tile index -> (site, tile) through a per-launch prefix table passed by value (BatchTab), the diagonal-block
// kernel runs one CTA per site, the memory-bound helpers take the site from the grid.  Per-site arrays are slabs
// [site][ld][ld] / [site][ld] with one 3-D tensor map per matrix (site = third TMA coordinate).
//
// Sites of different size are END-ALIGNED: with NB = max nb and panels of pw block columns, site i starts at global block
// step off_i = floor((NB - nb_i) / pw) pw, so that at every global step all active sites have (almost) the same trailing
// matrix left, all finish together, and the latency-bound chain (diagonal block -> panel solve -> in-panel update) is paid
// once per global step for the whole batch instead of once per site.  Panels of a site still start at its local block 0, so
// every tile sees exactly the operations (and operand order) of the single-site schedule: results are bit-identical.
//
// Included at the end of dgp_api.cu (uses its launch helpers).
#pragma once

struct dgp_batch_s {
  int device = 0, sms = 148;
  cudaStream_t stream = nullptr, stream_hi = nullptr, stream_lo = nullptr, stream_t2 = nullptr;
  bool own_stream = false;
  bool eager_inv = true;     // merges of the inverse launched as the factorisation passes them (DGP_EAGER_INV=0: after it)
  int eager_lag = 0;
  int strip_blocks = 8;      // column strips of the trailing updates on two streams (DGP_STRIP_BLOCKS; 0: one stream)
  cudaEvent_t ev_lo[2] = {nullptr, nullptr};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> evs;
  int panel_blocks = 4;
  bool pdl = true, chain_half = true, timing = false;
  bool pregen = true;        // standalone covariance generator ahead of the factorisation (DGP_PREGEN=0: first-touch generation)
  bool inpanel_left = false; // in-panel updates left-looking (DGP_INPANEL_LEFT=1; default: right-looking rank-128 updates, like the single-site engine -- the two must agree for bit-identical results)
  int max_sites = 0, max_pad = 0, max_n = 0;
  int G = 0, NB = 0;
  long long ld = 0;
  int n[DGP_BATCH_MAX] = {0}, nb[DGP_BATCH_MAX] = {0}, off[DGP_BATCH_MAX] = {0}, order[DGP_BATCH_MAX] = {0};
  bool have_train = false, pending = false;
  dgp_spec spec, user_spec;
  SiteDims sd;
  double *A = nullptr, *L = nullptr, *U = nullptr;
  double *X = nullptr, *y = nullptr, *noise = nullptr, *Xw = nullptr, *r = nullptr, *z = nullptr, *alpha = nullptr;
  double *theta = nullptr, *scal = nullptr, *gpart = nullptr, *zpart = nullptr, *jitv = nullptr;
  double *h_theta = nullptr, *h_scal = nullptr, *h_jit = nullptr;  // pinned
  CUtensorMap tmA, tmL, tmU;
  long long launches = 0;
  double last_ms[4] = {0, 0, 0, 0};
  std::string err;
};
typedef struct dgp_batch_s* dgp_batch_t;

static int make_map3(dgp_batch_t h, CUtensorMap* m, double* base, long long ld, int sites) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) DGP_FAIL(h, -3, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)ld, (cuuint64_t)sites};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * (cuuint64_t)ld * 8};
  cuuint32_t box[3] = {BK, 64, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DGP_FAIL(h, -3, "cuTensorMapEncodeTiled (3-D) failed (%d) ld=%lld sites=%d", (int)r, ld, sites);
  return 0;
}

namespace {
struct TabBuilder {
  BatchTab t;
  int total = 0;
  bool any_first = false;
  explicit TabBuilder(dgp_batch_t b) {
    memset(&t, 0, sizeof(t));
    t.ld = b->ld;
    t.jitv = b->jitv;
  }
  void add(int site, int step, int nb, int n, int a0, int a1, int a2, int ntiles) {
    if (ntiles <= 0) return;
    BatchEnt& e = t.e[t.count++];
    e.tile0 = total; e.site = site; e.step = step; e.nb = nb; e.n = n; e.aux0 = a0; e.aux1 = a1; e.aux2 = a2;
    total += ntiles;
  }
};

GemmArgs b_args(dgp_batch_t b, int mode, double* C, double sign, int ntiles) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.mode = mode; g.nb = b->NB; g.n = 0;
  g.ntiles = ntiles; g.C = C; g.ldc = b->ld; g.sign = sign;
  g.Xw = b->Xw; g.noise = b->noise; g.theta = b->theta; g.alpha = b->alpha; g.part = b->gpart;
  g.raster = raster_on();
  return g;
}

// one rank-(128 kb) update launch over the sites of `tb` (M_TRAIL_COL / M_TRAIL entries); half_limit: largest tile count
// that still runs on 64-row half tiles (0: never)
int b_launch_trail(dgp_batch_t b, const TabBuilder& tb, int mode, cudaStream_t st, int half_limit, bool pdl) {
  if (tb.total <= 0) return 0;
  GemmArgs g = b_args(b, mode, b->A, -1.0, tb.total);
  if (tb.any_first) return launch_gemm<INIT_COV, EPI_STORE>(b, b->tmL, b->tmL, g, st, pdl, &tb.t);
  if (b->chain_half && tb.total <= half_limit) return launch_gemm<INIT_LOAD, EPI_STORE, 4>(b, b->tmL, b->tmL, g, st, pdl, &tb.t);
  return launch_gemm<INIT_LOAD, EPI_STORE>(b, b->tmL, b->tmL, g, st, pdl, &tb.t);
}

int b_ensure_events(dgp_batch_t b, size_t count) {
  while (b->evs.size() < count) {
    cudaEvent_t e;
    CK(b, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    b->evs.push_back(e);
  }
  return 0;
}

// Two-level right-looking Cholesky with look-ahead (potrf_core of dgp_api.cu) over all sites, end-aligned.
struct BInvProgress;
int b_trtri_advance(dgp_batch_t b, BInvProgress& ip, int Fg);
int b_eager(dgp_batch_t b, BInvProgress* ip, cudaEvent_t ev_panel, int pe);

int b_potrf(dgp_batch_t b, BInvProgress* eager) {
  const int G = b->G, NB = b->NB, pw = b->panel_blocks;
  const long long ld = b->ld;
  cudaStream_t T = b->stream, P = b->stream_hi;
  const int npanels = (NB + pw - 1) / pw;
  int rc;
  const int sw = (b->strip_blocks + pw - 1) / pw * pw;
  cudaStream_t T2 = (sw > 0 && b->stream_t2 != nullptr) ? b->stream_t2 : nullptr;
  bool any2 = false;
  if ((rc = b_ensure_events(b, 3 * (size_t)npanels + 3))) return rc;
  auto ev_panel = [&](int p) { return b->evs[3 * p]; };
  auto ev_cols = [&](int p) { return b->evs[3 * p + 1]; };
  auto ev_col0 = [&](int p) { return b->evs[3 * p + 2]; };
  k_cov_col0_b<<<dim3(4 * NB, 1, G), 256, 0, T>>>(b->spec, b->sd, b->theta, b->Xw, b->noise, b->jitv, b->A);
  b->launches++;
  CK(b, cudaGetLastError());
  CK(b, cudaEventRecord(ev_cols(0), T));
  CK(b, cudaStreamWaitEvent(P, ev_cols(0), 0));
  // standalone generator (potrf_core's DGP_PREGEN schedule): the sites with more than DGP_PREGEN_MIN_NB block columns get their
  // whole lower triangle written on the low-priority stream while the first diagonal blocks are factored, and their first
  // updates read it like any later one.  The same sites by the same rule as the single-site engine: bit-identical results.
  bool pre[DGP_BATCH_MAX];
  bool any_pre = false, wP = false, wT = false, w2 = false;
  for (int i = 0; i < G; i++) {
    pre[i] = b->pregen && b->nb[i] > DGP_PREGEN_MIN_NB;
    any_pre |= pre[i];
  }
  cudaEvent_t ev_gen = b->evs[3 * (size_t)npanels + 1];
  if (any_pre) {
    CK(b, cudaStreamWaitEvent(b->stream_lo, ev_cols(0), 0));
    k_cov_lower_b<<<dim3(4 * NB, NB - 1, G), 256, 0, b->stream_lo>>>(b->spec, b->sd, b->theta, b->Xw, b->noise, b->jitv, b->A,
                                                                            DGP_PREGEN_MIN_NB);
    b->launches++;
    CK(b, cudaGetLastError());
    CK(b, cudaEventRecord(ev_gen, b->stream_lo));
  }
  const int slots = 2 * b->sms;
  for (int p = 0; p < npanels; p++) {
    const int pb = p * pw, pe = (pb + pw < NB) ? pb + pw : NB;
    if (p > 0) CK(b, cudaStreamWaitEvent(P, ev_col0(p), 0));
    // ---- the latency-bound chain of the panel, stream P
    for (int t = pb; t < pe; t++) {
      P2Batch pk;
      memset(&pk, 0, sizeof(pk));
      pk.slab = ld * ld;
      TabBuilder trsm(b), inp(b);
      for (int k = 0; k < G; k++) {
        const int i = b->order[k], s = t - b->off[i], nbi = b->nb[i];
        if (s < 0 || s >= nbi) continue;
        pk.site[pk.count] = (short)i; pk.step[pk.count] = (short)s; pk.count++;
        const int m = nbi - s - 1;
        trsm.add(i, s, nbi, b->n[i], 0, 1, 0, 2 * m);
        const int lpe = (pe - b->off[i] < nbi) ? pe - b->off[i] : nbi;
        if (s + 1 < lpe && b->inpanel_left) {
          // left-looking inside the panel (option): block column s + 1 takes the panel's columns [lpb, s] in ONE
          // rank-(128 (s+1-lpb)) update (read and written once per panel instead of once per earlier column).  Since the
          // tile engine adds a tile's old value in its epilogue (out = old + sign * sum, sums started at zero) the two forms
          // round differently, and with the C tile prefetched they cost the same (measured): the default is right-looking
          // in both engines, which keeps batch and single-site results bit-identical.
          const int lpb = pb - b->off[i];
          const bool first = (lpb == 0) && !pre[i];   // columns of a site's first panel are generated here
          inp.add(i, lpb, nbi, b->n[i], (s + 1) | (1 << 16), s + 1 - lpb, first ? 0 : 1, m * 2);
          inp.any_first |= first && m > 0;
        } else if (s + 1 < lpe) {
          const int w = lpe - s - 1;
          const bool first = (s == 0) && !pre[i];
          inp.add(i, s, nbi, b->n[i], (s + 1) | (w << 16), 1, first ? 0 : 1, m * 2 * w);
          inp.any_first |= first && m > 0;
        }
      }
      if (pk.count == 0) continue;
      // dependent launch behind the in-panel update of the previous block column (same stream, nothing in between)
      CK(b, launch_ex(k_potf2_v2, pk.count, P2_THREADS, (size_t)P2_SMEM, P, b->pdl && t > pb, (const double*)b->A, b->L, b->U, ld,
                      (double*)nullptr, b->scal, 0, b->A, 0, pk));
      b->launches++;
      if (trsm.total > 0) {  // panel solve: L[i, s] = A[i, s] T_ss', T_ss read from the diagonal block of the work matrix
        GemmArgs g = b_args(b, M_TRSM, b->L, 1.0, trsm.total);
        if (b->chain_half && 2 * trsm.total <= slots) { if ((rc = launch_gemm<INIT_ZERO, EPI_STORE, 4>(b, b->tmA, b->tmA, g, P, b->pdl, &trsm.t))) return rc; }
        else if ((rc = launch_gemm<INIT_ZERO, EPI_STORE>(b, b->tmA, b->tmA, g, P, b->pdl, &trsm.t))) return rc;
      }
      if (inp.total > 0) {   // rank-128 update of the panel's own remaining columns
        const bool waited = (t == pb && p > 0);
        if (waited) CK(b, cudaStreamWaitEvent(P, ev_cols(p), 0));
        const bool gen_wait = any_pre && !wP;
        if (gen_wait) { wP = true; CK(b, cudaStreamWaitEvent(P, ev_gen, 0)); }
        if ((rc = b_launch_trail(b, inp, M_TRAIL_COL, P, b->sms, b->pdl && !waited && !gen_wait))) return rc;
      }
    }
    CK(b, cudaEventRecord(ev_panel(p), P));
    CK(b, cudaStreamWaitEvent(T, ev_panel(p), 0));
    if (any_pre && !wT) { wT = true; CK(b, cudaStreamWaitEvent(T, ev_gen, 0)); }
    if (eager != nullptr && pe < NB && (rc = b_eager(b, eager, ev_panel(p), pe))) return rc;
    // ---- rank-(128 pw) update right of the panel.  Column strips on two streams (potrf_core of dgp_api.cu): fixed strips
    // of `sw` block columns in the end-aligned (global) column numbering, strip j always on stream j mod 2; the next panel's
    // columns are the head of their strip: first column | its other columns | the rest of the strip.
    if (T2 != nullptr) {
      const int j0 = pe / sw;
      bool used2 = false;
      auto strip_stream = [&](int j) -> cudaStream_t {
        if ((j & 1) == 0) return T;
        if (!used2) { used2 = true; cudaStreamWaitEvent(T2, ev_panel(p), 0); }
        if (any_pre && !w2) { w2 = true; cudaStreamWaitEvent(T2, ev_gen, 0); }
        return T2;
      };
      cudaStream_t S0 = strip_stream(j0);
      TabBuilder ta(b), tb2(b), tr(b);
      int jmax = j0;
      for (int k = 0; k < G; k++) {
        const int i = b->order[k], nbi = b->nb[i], lpb = pb - b->off[i];
        if (lpb < 0 || lpb >= nbi) continue;
        const int lpe = (pe - b->off[i] < nbi) ? pe - b->off[i] : nbi;
        if (lpe >= nbi) continue;
        const int kb = lpe - lpb, ne = (lpe + pw < nbi) ? lpe + pw : nbi, w = ne - lpe, m = nbi - lpe;
        const bool first = (lpb == 0) && !pre[i];
        const int a2 = first ? 0 : 1;
        ta.add(i, lpb, nbi, b->n[i], lpe | (1 << 16), kb, a2, m * 2);
        ta.any_first |= first;
        if (w > 1) { tb2.add(i, lpb, nbi, b->n[i], (lpe + 1) | ((w - 1) << 16), kb, a2, (m - 1) * 2 * (w - 1)); tb2.any_first |= first; }
        const int e0 = ((j0 + 1) * sw - b->off[i] < nbi) ? (j0 + 1) * sw - b->off[i] : nbi;   // local end of the near strip
        if (ne < e0) { tr.add(i, lpb, nbi, b->n[i], ne | ((e0 - ne) << 16), kb, a2, (nbi - ne) * 2 * (e0 - ne)); tr.any_first |= first; }
        const int jl = (b->off[i] + nbi - 1) / sw;   // last global strip this site reaches
        if (jl > jmax) jmax = jl;
      }
      if ((rc = b_launch_trail(b, ta, M_TRAIL_COL, S0, slots, false))) return rc;
      if (p + 1 < npanels) CK(b, cudaEventRecord(ev_col0(p + 1), S0));
      if ((rc = b_launch_trail(b, tb2, M_TRAIL_COL, S0, slots, false))) return rc;
      if (p + 1 < npanels) CK(b, cudaEventRecord(ev_cols(p + 1), S0));
      if ((rc = b_launch_trail(b, tr, M_TRAIL_COL, S0, slots, false))) return rc;
      for (int j = j0 + 1; j <= jmax; j++) {
        TabBuilder ts(b);
        for (int k = 0; k < G; k++) {
          const int i = b->order[k], nbi = b->nb[i], lpb = pb - b->off[i];
          if (lpb < 0 || lpb >= nbi) continue;
          const int lpe = (pe - b->off[i] < nbi) ? pe - b->off[i] : nbi;
          const int o = j * sw - b->off[i];
          if (lpe >= nbi || o >= nbi) continue;
          const int wj = (o + sw < nbi) ? sw : nbi - o;
          const bool first = (lpb == 0) && !pre[i];
          ts.add(i, lpb, nbi, b->n[i], o | (wj << 16), lpe - lpb, first ? 0 : 1, (nbi - o) * 2 * wj);
          ts.any_first |= first;
        }
        if (ts.total > 0 && (rc = b_launch_trail(b, ts, M_TRAIL_COL, strip_stream(j), slots, false))) return rc;
      }
      if (used2) any2 = true;
      continue;
    }
    // one stream: next panel's first column | its other columns | the rest
    TabBuilder ta(b), tb2(b), tc(b);
    for (int k = 0; k < G; k++) {
      const int i = b->order[k], nbi = b->nb[i], lpb = pb - b->off[i];
      if (lpb < 0 || lpb >= nbi) continue;
      const int lpe = (pe - b->off[i] < nbi) ? pe - b->off[i] : nbi;
      if (lpe >= nbi) continue;
      const int kb = lpe - lpb, ne = (lpe + pw < nbi) ? lpe + pw : nbi, w = ne - lpe, m = nbi - lpe, m2 = nbi - ne;
      const bool first = (lpb == 0) && !pre[i];
      const int a2 = first ? 0 : 1;
      ta.add(i, lpb, nbi, b->n[i], lpe | (1 << 16), kb, a2, m * 2);
      ta.any_first |= first;
      if (w > 1) { tb2.add(i, lpb, nbi, b->n[i], (lpe + 1) | ((w - 1) << 16), kb, a2, (m - 1) * 2 * (w - 1)); tb2.any_first |= first; }
      if (m2 > 0) { tc.add(i, lpb, nbi, b->n[i], ne, kb, a2, m2 * (m2 + 1)); tc.any_first |= first; }
    }
    if ((rc = b_launch_trail(b, ta, M_TRAIL_COL, T, slots, false))) return rc;
    if (p + 1 < npanels) CK(b, cudaEventRecord(ev_col0(p + 1), T));
    if ((rc = b_launch_trail(b, tb2, M_TRAIL_COL, T, slots, false))) return rc;
    if (p + 1 < npanels) CK(b, cudaEventRecord(ev_cols(p + 1), T));
    if ((rc = b_launch_trail(b, tc, M_TRAIL, T, slots, false))) return rc;
  }
  if (any2) {  // the caller continues on T
    CK(b, cudaEventRecord(b->evs[3 * (size_t)npanels], T2));
    CK(b, cudaStreamWaitEvent(T, b->evs[3 * (size_t)npanels], 0));
  }
  return 0;
}

// U = L^-T by recursive doubling, all sites per launch, stream W.  b_trtri_advance(Fg): every merge whose block range lies
// inside the block columns that are final once the global step Fg is reached (site i: its first Fg - off_i columns) and has
// not been launched yet (trtri_advance of dgp_api.cu); Fg >= NB: everything that is left.
struct BInvProgress {
  short done[16][DGP_BATCH_MAX];
  cudaStream_t W;
  int lag;
};

int b_trtri_advance(dgp_batch_t b, BInvProgress& ip, int Fg) {
  const int G = b->G, NB = b->NB;
  int rc, lv = 0;
  for (int hb = 1; hb < NB; hb *= 2, lv++) {
    TabBuilder tb(b);
    PairRange prg;
    memset(&prg, 0, sizeof(prg));
    int tmax = 0;
    for (int k = 0; k < G; k++) {
      const int i = b->order[k], nbi = b->nb[i];
      if (nbi <= hb) continue;
      const int npairs = (nbi - hb + 2 * hb - 1) / (2 * hb);
      const int F = Fg >= NB ? nbi : Fg - b->off[i];
      const int avail = F >= nbi ? npairs : (F > 0 ? F / (2 * hb) : 0);
      const int pr0 = ip.done[lv][i], cnt = avail - pr0;
      if (cnt <= 0) continue;
      tb.add(i, pr0, nbi, b->n[i], hb, cnt, 0, cnt * hb * 2 * hb);
      if (2 * hb < nbi) {
        prg.pr0[i] = (short)pr0; prg.cnt[i] = (short)cnt;
        if (cnt * 16 * hb * hb > tmax) tmax = cnt * 16 * hb * hb;
      }
      ip.done[lv][i] = (short)avail;
    }
    if (tb.total <= 0) continue;
    {
      GemmArgs g = b_args(b, M_INV_M, b->A, 1.0, tb.total);
      if ((rc = launch_gemm<INIT_ZERO, EPI_STORE>(b, b->tmU, b->tmL, g, ip.W, false, &tb.t))) return rc;
    }
    {
      GemmArgs g = b_args(b, M_INV_U, b->U, -1.0, tb.total);
      if ((rc = launch_gemm<INIT_ZERO, EPI_STORE>(b, b->tmA, b->tmA, g, ip.W, false, &tb.t))) return rc;
    }
    if (tmax > 0) {
      k_transpose_pairs_b<<<dim3(tmax, G), 256, 0, ip.W>>>(b->sd, prg, b->U, b->A, hb);
      b->launches++;
      CK(b, cudaGetLastError());
    }
  }
  return 0;
}

int b_eager(dgp_batch_t b, BInvProgress* ip, cudaEvent_t ev_panel, int pe) {
  const int F = pe - ip->lag;
  if (F < 2 || (F & 7) != 0) return 0;
  CK(b, cudaStreamWaitEvent(ip->W, ev_panel, 0));
  return b_trtri_advance(b, *ip, F);
}

// the rest of the inverse, then z = U'r, alpha = U z
int b_trtri(dgp_batch_t b, BInvProgress& ip) {
  const int G = b->G, NB = b->NB;
  cudaStream_t W = ip.W;
  int rc;
  if ((rc = b_trtri_advance(b, ip, NB))) return rc;
  k_upperT_gemv_part_b<<<dim3(NB, NB, G), 256, 0, W>>>(b->sd, b->U, b->r, b->zpart);
  k_upperT_gemv_sum_b<<<dim3((unsigned)((b->ld + 255) / 256), G), 256, 0, W>>>(b->sd, b->zpart, b->z);
  k_upper_gemv_b<<<dim3((unsigned)(b->ld / 8), G), 256, 0, W>>>(b->sd, b->U, b->z, b->alpha);
  b->launches += 3;
  CK(b, cudaGetLastError());
  return 0;
}

// LAUUM -> Ky^-1 (lower tiles, over the dead T) and the gradient contraction W (.) dK/dtheta; all sites per launch.
int b_lauum_grad(dgp_batch_t b, cudaStream_t W) {
  TabBuilder tb(b);
  for (int k = 0; k < b->G; k++) {
    const int i = b->order[k];
    tb.add(i, 0, b->nb[i], b->n[i], 0, 0, 0, lauum_slots(b->nb[i], raster_on()));
  }
  GemmArgs g = b_args(b, M_LAUUM, b->A, 1.0, tb.total);
  int rc = launch_gemm<INIT_ZERO, EPI_STORE>(b, b->tmU, b->tmU, g, W, false, &tb.t);
  if (rc) return rc;
  k_grad_contract_b<<<b->sd.tile0[b->G], 128, 0, W>>>(b->spec, b->sd, b->theta, b->Xw, b->alpha, b->A, b->gpart);
  b->launches++;
  CK(b, cudaGetLastError());
  return 0;
}

int b_enqueue(dgp_batch_t b) {
  int rc;
  const int G = b->G;
  cudaStream_t T = b->stream, W = b->stream_lo ? b->stream_lo : b->stream;
  if (b->timing) CK(b, cudaEventRecord(b->ev[0], T));
  CK(b, cudaMemcpyAsync(b->theta, b->h_theta, sizeof(double) * DGP_MAX_THETA * G, cudaMemcpyHostToDevice, T));
  CK(b, cudaMemcpyAsync(b->jitv, b->h_jit, sizeof(double) * G, cudaMemcpyHostToDevice, T));
  k_features_b<<<dim3((unsigned)((b->ld + 255) / 256), G), 256, 0, T>>>(b->spec, b->sd, b->theta, b->X, b->y, b->Xw, b->r, b->scal);
  b->launches++;
  CK(b, cudaGetLastError());
  BInvProgress ip;
  memset(&ip, 0, sizeof(ip));
  ip.W = W; ip.lag = b->eager_lag;
  if ((rc = b_potrf(b, (b->eager_inv && W != T) ? &ip : nullptr))) return rc;
  if (b->timing) CK(b, cudaEventRecord(b->ev[1], T));
  if (W != T) {
    CK(b, cudaEventRecord(b->ev_lo[0], T));
    CK(b, cudaStreamWaitEvent(W, b->ev_lo[0], 0));
  }
  if ((rc = b_trtri(b, ip))) return rc;
  if (b->timing) CK(b, cudaEventRecord(b->ev[2], W));
  if ((rc = b_lauum_grad(b, W))) return rc;
  if (b->timing) CK(b, cudaEventRecord(b->ev[3], W));
  if (W != T) {
    CK(b, cudaEventRecord(b->ev_lo[1], W));
    CK(b, cudaStreamWaitEvent(T, b->ev_lo[1], 0));
  }
  k_finish_b<<<dim3(b->spec.ntheta + 1, G), 256, 0, T>>>(b->spec, b->sd, b->theta, b->gpart, b->z, b->alpha, b->X, b->scal);
  b->launches++;
  CK(b, cudaGetLastError());
  CK(b, cudaMemcpyAsync(b->h_scal, b->scal, sizeof(double) * SC_SIZE * G, cudaMemcpyDeviceToHost, T));
  if (b->timing) CK(b, cudaEventRecord(b->ev[4], T));
  return 0;
}
}  // namespace

extern "C" {

size_t dgp_batch_workspace_bytes(int max_sites, int max_n) {
  const size_t np = round_up(max_n > 0 ? max_n : 128, 128), S = max_sites > 0 ? max_sites : 1, nbm = np / 128;
  return S * (3 * np * np + np * (DGP_MAX_COLS + DGP_XS + 6) + nbm * (nbm + 1) * DGP_MAX_THETA + nbm * np + DGP_MAX_THETA + SC_SIZE + 1) * 8;
}

int dgp_batch_destroy(dgp_batch b) {
  if (!b) return 0;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  if (b->stream_hi) { cudaStreamSynchronize(b->stream_hi); cudaStreamDestroy(b->stream_hi); }
  if (b->stream_lo) { cudaStreamSynchronize(b->stream_lo); cudaStreamDestroy(b->stream_lo); }
  if (b->stream_t2) { cudaStreamSynchronize(b->stream_t2); cudaStreamDestroy(b->stream_t2); }
  for (int i = 0; i < 2; i++) if (b->ev_lo[i]) cudaEventDestroy(b->ev_lo[i]);
  for (int i = 0; i < 5; i++) if (b->ev[i]) cudaEventDestroy(b->ev[i]);
  for (cudaEvent_t e : b->evs) cudaEventDestroy(e);
  double* bufs[] = {b->A, b->L, b->U, b->X, b->y, b->noise, b->Xw, b->r, b->z, b->alpha, b->theta, b->scal, b->gpart, b->zpart, b->jitv};
  for (double* p : bufs) if (p) cudaFree(p);
  if (b->h_theta) cudaFreeHost(b->h_theta);
  if (b->h_scal) cudaFreeHost(b->h_scal);
  if (b->h_jit) cudaFreeHost(b->h_jit);
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  delete b;
  return 0;
}

int dgp_batch_create(dgp_batch* out, int device, int max_sites, int max_n, void* stream) {
  dgp_batch_t nil = nullptr;
  if (!out || max_n <= 0 || max_sites < 1 || max_sites > DGP_BATCH_MAX)
    DGP_FAIL(nil, -1, "dgp_batch_create: bad arguments (1 <= max_sites <= %d)", DGP_BATCH_MAX);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    DGP_FAIL(nil, -2, "dgp_batch_create: no CUDA device (this engine has no CPU fallback)");
  if (cudaSetDevice(device) != cudaSuccess) DGP_FAIL(nil, -2, "dgp_batch_create: cannot select device %d", device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) DGP_FAIL(nil, -2, "dgp_batch_create: libdgp is built for sm_100a (B200) only");
  dgp_batch_t b = new dgp_batch_s();
  b->device = device;
  b->sms = prop.multiProcessorCount;
  b->max_sites = max_sites; b->max_n = max_n; b->max_pad = round_up(max_n, 128);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (stream) b->stream = (cudaStream_t)stream;
  else {
    cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, (lo + hi) / 2);
    b->own_stream = true;
  }
  cudaStreamCreateWithPriority(&b->stream_lo, cudaStreamNonBlocking, lo);
  const char* ei = getenv("DGP_EAGER_INV");
  if (ei) b->eager_inv = atoi(ei) != 0;
  const char* el = getenv("DGP_EAGER_LAG");
  if (el && atoi(el) >= 0) b->eager_lag = atoi(el) / 8 * 8;
  cudaStreamCreateWithPriority(&b->stream_hi, cudaStreamNonBlocking, hi);
  {
    int pr = (lo + hi) / 2;
    if (cudaStreamGetPriority(b->stream, &pr) != cudaSuccess) { cudaGetLastError(); pr = (lo + hi) / 2; }
    cudaStreamCreateWithPriority(&b->stream_t2, cudaStreamNonBlocking, pr);
  }
  const char* sbk = getenv("DGP_STRIP_BLOCKS");
  if (sbk && atoi(sbk) >= 0 && atoi(sbk) <= 4096) b->strip_blocks = atoi(sbk);
  const char* pd = getenv("DGP_PDL");
  if (pd) b->pdl = atoi(pd) != 0;
  const char* ch = getenv("DGP_CHAIN_HALF");
  if (ch) b->chain_half = atoi(ch) != 0;
  const char* pbk = getenv("DGP_PANEL_BLOCKS");
  if (pbk && atoi(pbk) >= 1 && atoi(pbk) <= 64) b->panel_blocks = atoi(pbk);
  const char* pg = getenv("DGP_PREGEN");
  if (pg) b->pregen = atoi(pg) != 0;
  const char* il = getenv("DGP_INPANEL_LEFT");
  if (il) b->inpanel_left = atoi(il) != 0;
  const size_t np = b->max_pad, S = max_sites, nbm = np / 128;
  cudaError_t r = cudaSuccess;
  auto A = [&](double** p, size_t count) { if (r == cudaSuccess) r = cudaMalloc((void**)p, count * sizeof(double)); };
  A(&b->A, S * np * np); A(&b->L, S * np * np); A(&b->U, S * np * np);
  A(&b->X, S * np * DGP_MAX_COLS); A(&b->y, S * np); A(&b->noise, S * np); A(&b->Xw, S * np * DGP_XS);
  A(&b->r, S * np); A(&b->z, S * np); A(&b->alpha, S * np);
  A(&b->theta, S * DGP_MAX_THETA); A(&b->scal, S * SC_SIZE); A(&b->gpart, S * nbm * (nbm + 1) * DGP_MAX_THETA);
  A(&b->zpart, S * nbm * np); A(&b->jitv, S);
  if (r == cudaSuccess) r = cudaMallocHost((void**)&b->h_theta, S * DGP_MAX_THETA * sizeof(double));
  if (r == cudaSuccess) r = cudaMallocHost((void**)&b->h_scal, S * SC_SIZE * sizeof(double));
  if (r == cudaSuccess) r = cudaMallocHost((void**)&b->h_jit, S * sizeof(double));
  for (int i = 0; i < 5 && r == cudaSuccess; i++) r = cudaEventCreate(&b->ev[i]);
  for (int i = 0; i < 2 && r == cudaSuccess; i++) r = cudaEventCreateWithFlags(&b->ev_lo[i], cudaEventDisableTiming);
  if (r != cudaSuccess) {
    g_create_error = std::string("dgp_batch_create: allocation failed: ") + cudaGetErrorString(r);
    cudaGetLastError();
    dgp_batch_destroy(b);
    return -2;
  }
  cudaFuncSetAttribute(k_potf2_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
  *out = b;
  return 0;
}

const char* dgp_batch_last_error(dgp_batch b) { return b ? b->err.c_str() : g_create_error.c_str(); }

int dgp_batch_set_train(dgp_batch b, const dgp_spec* spec, int nsites, const int* n, const double* const* X,
                        const double* const* y, const double* const* noise) {
  if (!b) return -1;
  if (!spec || !n || !X || !y || !noise || nsites < 1 || nsites > b->max_sites)
    DGP_FAIL(b, -1, "dgp_batch_set_train: bad arguments (nsites=%d, max_sites=%d)", nsites, b->max_sites);
  int rc = check_spec(b, spec);
  if (rc) return rc;
  for (int i = 0; i < nsites; i++)
    if (n[i] < 1 || n[i] > b->max_n || !X[i] || !y[i] || !noise[i])
      DGP_FAIL(b, -1, "dgp_batch_set_train: site %d: n=%d (max_n=%d) or NULL data", i, n[i], b->max_n);
  CK(b, cudaSetDevice(b->device));
  if (b->pending) { CK(b, cudaStreamSynchronize(b->stream)); b->pending = false; }
  b->user_spec = *spec;
  b->spec = *spec;
  augment_spec(&b->spec);
  b->G = nsites;
  int npmax = 0;
  for (int i = 0; i < nsites; i++) {
    b->n[i] = n[i];
    b->nb[i] = round_up(n[i], 128) / 128;
    if (b->nb[i] * 128 > npmax) npmax = b->nb[i] * 128;
  }
  b->ld = npmax;
  b->NB = npmax / 128;
  const int pw = b->panel_blocks;
  memset(&b->sd, 0, sizeof(b->sd));
  b->sd.count = nsites; b->sd.nbmax = b->NB; b->sd.ld = b->ld;
  for (int i = 0; i < nsites; i++) {
    b->off[i] = ((b->NB - b->nb[i]) / pw) * pw;
    b->sd.n[i] = b->n[i]; b->sd.nb[i] = b->nb[i];
    b->sd.tile0[i + 1] = b->sd.tile0[i] + b->nb[i] * (b->nb[i] + 1);
    b->order[i] = i;
  }
  // launch tables list the sites largest first (longest tiles of a launch start first)
  for (int i = 1; i < nsites; i++)
    for (int k = i; k > 0 && b->nb[b->order[k]] > b->nb[b->order[k - 1]]; k--) std::swap(b->order[k], b->order[k - 1]);
  const size_t ld = (size_t)b->ld, slab = ld * ld;
  cudaStream_t st = b->stream;
  CK(b, cudaMemsetAsync(b->L, 0, slab * nsites * 8, st));
  CK(b, cudaMemsetAsync(b->U, 0, slab * nsites * 8, st));
  CK(b, cudaMemsetAsync(b->noise, 0, ld * nsites * 8, st));
  CK(b, cudaMemsetAsync(b->X, 0, ld * nsites * DGP_MAX_COLS * 8, st));
  CK(b, cudaMemsetAsync(b->y, 0, ld * nsites * 8, st));
  for (int i = 0; i < nsites; i++) {
    CK(b, cudaMemcpyAsync(b->X + (size_t)i * ld * DGP_MAX_COLS, X[i], (size_t)n[i] * spec->ndim * 8, cudaMemcpyHostToDevice, st));
    CK(b, cudaMemcpyAsync(b->y + (size_t)i * ld, y[i], (size_t)n[i] * 8, cudaMemcpyHostToDevice, st));
    CK(b, cudaMemcpyAsync(b->noise + (size_t)i * ld, noise[i], (size_t)n[i] * 8, cudaMemcpyHostToDevice, st));
  }
  CK(b, cudaStreamSynchronize(st));
  if ((rc = make_map3(b, &b->tmA, b->A, b->ld, nsites))) return rc;
  if ((rc = make_map3(b, &b->tmL, b->L, b->ld, nsites))) return rc;
  if ((rc = make_map3(b, &b->tmU, b->U, b->ld, nsites))) return rc;
  b->have_train = true;
  return 0;
}

int dgp_batch_nlml_grad_launch(dgp_batch b, const double* theta, const double* jitter) {
  if (!b) return -1;
  if (!b->have_train) DGP_FAIL(b, -1, "no training data: call dgp_batch_set_train first");
  if (!theta) DGP_FAIL(b, -1, "theta is NULL");
  if (b->pending) DGP_FAIL(b, -1, "an evaluation is already in flight: call dgp_batch_nlml_grad_wait first");
  CK(b, cudaSetDevice(b->device));
  const int P = b->spec.ntheta;
  memset(b->h_theta, 0, sizeof(double) * DGP_MAX_THETA * b->G);
  for (int i = 0; i < b->G; i++) {
    memcpy(b->h_theta + (size_t)i * DGP_MAX_THETA, theta + (size_t)i * P, sizeof(double) * P);
    b->h_jit[i] = jitter ? jitter[i] : 0.0;
  }
  int rc = b_enqueue(b);
  if (rc) return rc;
  b->pending = true;
  return 0;
}

int dgp_batch_nlml_grad_ready(dgp_batch b) {
  if (!b) return -1;
  if (!b->pending) DGP_FAIL(b, -1, "no evaluation in flight");
  CK(b, cudaSetDevice(b->device));
  const cudaError_t e = cudaStreamQuery(b->stream);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  DGP_FAIL(b, -2, "cudaStreamQuery failed: %s", cudaGetErrorString(e));
}

int dgp_batch_nlml_grad_wait(dgp_batch b, double* nlml_out, double* grad_out, int* info_out) {
  if (!b) return -1;
  if (!b->pending) DGP_FAIL(b, -1, "no evaluation in flight");
  CK(b, cudaSetDevice(b->device));
  CK(b, cudaStreamSynchronize(b->stream));
  b->pending = false;
  if (b->timing) {
    float ms;
    for (int i = 0; i < 4; i++) { cudaEventElapsedTime(&ms, b->ev[i], b->ev[i + 1]); b->last_ms[i] = ms; }
  }
  const int P = b->spec.ntheta;
  int bad = 0;
  for (int i = 0; i < b->G; i++) {
    const double* sc = b->h_scal + (size_t)i * SC_SIZE;
    if (nlml_out) nlml_out[i] = sc[SC_NLML];
    if (grad_out) memcpy(grad_out + (size_t)i * P, sc + SC_GRAD, sizeof(double) * P);
    const int info = (int)sc[SC_INFO];
    if (info_out) info_out[i] = info;
    if (info != 0) bad++;
  }
  return bad;
}

int dgp_batch_nlml_grad(dgp_batch b, const double* theta, const double* jitter, double* nlml_out, double* grad_out,
                        int* info_out) {
  int rc = dgp_batch_nlml_grad_launch(b, theta, jitter);
  if (rc) return rc;
  return dgp_batch_nlml_grad_wait(b, nlml_out, grad_out, info_out);
}

int dgp_batch_get_alpha(dgp_batch b, int site, double* alpha_out) {
  if (!b) return -1;
  if (!b->have_train || b->pending || site < 0 || site >= b->G || !alpha_out) DGP_FAIL(b, -1, "dgp_batch_get_alpha: bad arguments");
  CK(b, cudaSetDevice(b->device));
  CK(b, cudaMemcpyAsync(alpha_out, b->alpha + (size_t)site * b->ld, (size_t)b->n[site] * 8, cudaMemcpyDeviceToHost, b->stream));
  CK(b, cudaStreamSynchronize(b->stream));
  return 0;
}

long long dgp_batch_launch_count(dgp_batch b) { return b ? b->launches : 0; }
int dgp_batch_set_timing(dgp_batch b, int enable) { if (!b) return -1; b->timing = enable != 0; return 0; }
int dgp_batch_last_timing(dgp_batch b, double* ms4) {
  if (!b || !ms4) return -1;
  for (int i = 0; i < 4; i++) ms4[i] = b->last_ms[i];
  return 0;
}

}  // extern "C"
