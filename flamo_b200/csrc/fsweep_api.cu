// fsweep_api.cu — the C ABI declared in include/fsweep.h: plan validation / lowering to step
// tables, workspace sizing, and the forward / backward launches.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <type_traits>
#include <vector>

#include "fsweep_cta.cuh"

#ifndef FSWEEP_CTA_TC_DEFAULT
#define FSWEEP_CTA_TC_DEFAULT true  // FSWEEP_CTA_TC=0 selects the SIMT elimination instead of the tensor-core one
#endif
#include "fsweep_stream.cuh"
#include "fsweep_streamr.cuh"
#include "fsweep_streamw.cuh"
#include "fsweep_tpr.cuh"

using namespace fsweep;

namespace {

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool kind_is_diag(int k) {
  return k == FSWEEP_OP_PGAIN || k == FSWEEP_OP_PSOS || k == FSWEEP_OP_PDELAY || k == FSWEEP_OP_PTABLE;
}
bool kind_is_table(int k) { return k == FSWEEP_OP_TABLE || k == FSWEEP_OP_PTABLE; }
bool kind_is_leaf(int k) { return k >= FSWEEP_OP_GAIN && k <= FSWEEP_OP_PTABLE; }

int row_len_of(const fsweep_op_t& o) {
  switch (o.kind) {
    case FSWEEP_OP_GAIN:
    case FSWEEP_OP_DELAY:
      return o.n_in;
    case FSWEEP_OP_PGAIN:
    case FSWEEP_OP_PDELAY:
      return 1;
    case FSWEEP_OP_SOS:
      return o.n_sections * o.n_in * 16;
    case FSWEEP_OP_PSOS:
      return o.n_sections * 16;
    default:
      return 0;
  }
}

constexpr int MAX_GRID = 2048;
constexpr size_t LOSS_PARTIAL_BYTES = (size_t)MAX_GRID * sizeof(double);  // fused criterion: per-block error sums
constexpr size_t SMEM_ACC_BUDGET = 64 * 1024;

}  // namespace

struct fsweep_plan {
  int dtype;
  int G;
  ProgK prog;         // template: everything but pointers and the gx-dependent flag
  int first_pre_rstep;  // index in rsteps of op 0's reverse step when it is a top-level op (-1 otherwise)
  int n_coeffs;
  std::vector<fsweep_op_t> leaf;  // leaf ops in slot order == ProgK::ops order
  bool any_global, any_acc;
  bool loop_fast = false;  // program matches the FDN-loop pattern of fsweep_loop.cuh
  LoopInfo loop;
  int tpb_np = 0;  // 4 / 8: additionally small enough for the thread-per-bin kernels of fsweep_tpb.cuh
  bool tpb_force = false;  // FSWEEP_FORCE_TPB=1 (tests): use them regardless of the bin count
  int stream_bps[2] = {0, 0};       // cached occupancy of the streaming kernels ...
  size_t stream_bps_smem[2] = {0, 0};  // ... for this dynamic shared memory size
  int streamr_bps = 0;                 // the same for the register-state forward kernel
  size_t streamr_bps_smem = 0;
  bool stream = false;  // TABLE-heavy program without recursion: streaming kernels, fsweep_stream.cuh
  StreamInfo sinfo;     // everything but tb / qc / threads (chosen per call from batch*cols)
  bool items = false;   // some op carries one coefficient set PER BATCH ITEM (fsweep_op_t::per_item): generic kernels, grid.y = item
  bool grad32 = false;  // FSWEEP_DT_GRAD32: float64 plan of a float32 model, deferred cascade gradients in float32 arithmetic
  bool cta = false;  // wide flagship shape (32 < N <= 64, float32): CTA-per-bin kernels, fsweep_cta.cuh
  bool cta_tc = false;  // ... with the tensor-core elimination (fsweep_tc.cuh) instead of the SIMT one
  int cta_blocks_per_sm[2] = {0, 0};
  int tpc_np = 0;  // 4 / 8: the flagship shape (N x 1 gain, loop width <= 8, 1 x N gain) -> compact kernels, fsweep_tpc.cuh
  bool tpc_force = false;  // FSWEEP_FORCE_TPC=1 (tests)
  // lazily filled launch geometry: [cc index 0:1, 1:4, 2:loop kernels][fwd, bwd]
  std::mutex mu;
  int blocks_per_sm[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  int num_sms = 0;
};

extern "C" int fsweep_version(void) { return FSWEEP_VERSION; }
extern "C" const char* fsweep_last_error(void) { return g_err; }
extern "C" int fsweep_last_launch_count(void) { return g_launches; }

extern "C" int fsweep_plan_create(const fsweep_op_t* ops, int n_ops, int64_t nfft, double alias_decay_db, int dtype,
                                  fsweep_plan_t** out) {
  if (!ops || !out || n_ops <= 0) return fail(FSWEEP_E_BADARG, "null or empty program");
  const bool grad32 = (dtype & FSWEEP_DT_GRAD32) != 0;
  dtype &= ~FSWEEP_DT_GRAD32;
  if (dtype != FSWEEP_C64 && dtype != FSWEEP_C128) return fail(FSWEEP_E_BADARG, "bad dtype %d", dtype);
  if (nfft < 2) return fail(FSWEEP_E_BADARG, "bad nfft %lld", (long long)nfft);

  // ---- split into pre | ff | fb | post and validate
  std::vector<int> pre, ff, fb, post;
  int rec = -1;
  for (int i = 0; i < n_ops;) {
    const fsweep_op_t& o = ops[i];
    if (o.kind == FSWEEP_OP_RECURSION) {
      if (rec >= 0) return fail(FSWEEP_E_UNSUPPORTED, "more than one RECURSION in one program (split the series)");
      if (o.n_ff < 1 || o.n_fb < 1)
        return fail(FSWEEP_E_BADARG, "RECURSION at %d: bad chain lengths (%d, %d)", i, o.n_ff, o.n_fb);
      if (i + 1 + o.n_ff + o.n_fb > n_ops) return fail(FSWEEP_E_BADARG, "RECURSION at %d: chains run past the end", i);
      rec = i;
      for (int j = 0; j < o.n_ff; ++j) ff.push_back(i + 1 + j);
      for (int j = 0; j < o.n_fb; ++j) fb.push_back(i + 1 + o.n_ff + j);
      i += 1 + o.n_ff + o.n_fb;
    } else {
      (rec < 0 ? pre : post).push_back(i);
      ++i;
    }
  }
  int width = 0;
  auto check_leaf = [&](int i) -> int {
    const fsweep_op_t& o = ops[i];
    if (!kind_is_leaf(o.kind))
      return fail(o.kind == FSWEEP_OP_RECURSION ? FSWEEP_E_UNSUPPORTED : FSWEEP_E_BADARG,
                  "op %d: kind %d not allowed here (nested recursion is unsupported)", i, o.kind);
    if (o.n_in < 1 || o.n_out < 1) return fail(FSWEEP_E_BADARG, "op %d: bad channel counts", i);
    if (kind_is_diag(o.kind) && o.n_in != o.n_out) return fail(FSWEEP_E_BADARG, "op %d: diagonal kind needs n_in == n_out", i);
    if ((o.kind == FSWEEP_OP_SOS || o.kind == FSWEEP_OP_PSOS) && o.n_sections < 1)
      return fail(FSWEEP_E_BADARG, "op %d: SOS needs n_sections >= 1", i);
    width = std::max(width, std::max(o.n_in, o.n_out));
    return FSWEEP_OK;
  };
  auto check_chain = [&](const std::vector<int>& c, const char* what) -> int {
    for (size_t j = 0; j < c.size(); ++j) {
      int r = check_leaf(c[j]);
      if (r) return r;
      if (j > 0 && ops[c[j]].n_in != ops[c[j - 1]].n_out)
        return fail(FSWEEP_E_BADARG, "%s chain: op %d has %d inputs but op %d has %d outputs", what, c[j], ops[c[j]].n_in,
                    c[j - 1], ops[c[j - 1]].n_out);
    }
    return FSWEEP_OK;
  };
  int r;
  if ((r = check_chain(pre, "series"))) return r;
  if ((r = check_chain(post, "series"))) return r;
  int rec_n = 0, rec_in = 0;
  if (rec >= 0) {
    if ((r = check_chain(ff, "feedforward"))) return r;
    if ((r = check_chain(fb, "feedback"))) return r;
    rec_in = ops[ff.front()].n_in;
    rec_n = ops[ff.back()].n_out;
    if (ops[fb.front()].n_in != rec_n || ops[fb.back()].n_out != rec_in)
      return fail(FSWEEP_E_BADARG, "recursion: feedforward is %d->%d but feedback is %d->%d", rec_in, rec_n,
                  ops[fb.front()].n_in, ops[fb.back()].n_out);
    if (!pre.empty() && ops[pre.back()].n_out != rec_in)
      return fail(FSWEEP_E_BADARG, "recursion input has %d channels but the preceding op outputs %d", rec_in,
                  ops[pre.back()].n_out);
    if (!post.empty() && ops[post.front()].n_in != rec_n)
      return fail(FSWEEP_E_BADARG, "recursion output has %d channels but the next op takes %d", rec_n,
                  ops[post.front()].n_in);
  }
  if (width > 64)
    return fail(FSWEEP_E_UNSUPPORTED, "channel width %d > 64: not supported by the register-resident sweep", width);
  const bool wide64 = width > 32 && dtype != FSWEEP_C64;  // float64 beyond 32 channels: the CTA-per-bin kernels only (below)
  const int n_leaf = (int)(pre.size() + ff.size() + fb.size() + post.size());
  if (n_leaf > MAX_OPS) return fail(FSWEEP_E_UNSUPPORTED, "program has %d ops (max %d): split the series", n_leaf, MAX_OPS);

  fsweep_plan* p = new (std::nothrow) fsweep_plan();
  if (!p) return fail(FSWEEP_E_BADARG, "out of host memory");
  p->dtype = dtype;
  p->grad32 = grad32 && dtype == FSWEEP_C128;
  int G = 1;
  while (G < width) G <<= 1;
  p->G = G;

  ProgK& P = p->prog;
  memset(&P, 0, sizeof(P));
  P.nfft = nfft;
  const double lng = -std::fabs(alias_decay_db) / (double)nfft / 20.0 * std::log(10.0);
  P.lng = lng;
  P.inv_nfft = 1.0 / (double)nfft;
  P.gm1 = std::expm1(lng);
  P.g2m1 = std::expm1(2.0 * lng);
  P.rec_n = rec_n;

  // kernel op order == coefficient slot order == order of appearance in the flat program
  std::vector<int> order;  // flat index of every leaf, ascending
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind != FSWEEP_OP_RECURSION) order.push_back(i);
  std::vector<int> slot_of(n_ops, -1);
  for (size_t s = 0; s < order.size(); ++s) slot_of[order[s]] = (int)s;
  p->n_coeffs = (int)order.size();
  P.n_ops = p->n_coeffs;
  const size_t real_sz = dtype == FSWEEP_C64 ? 4 : 8;

  int acc_per_lane = 0, acc_total = 0, h_total = 0;
  p->any_global = p->any_acc = false;
  for (int i : order) p->items = p->items || ops[i].per_item != 0;
  P.needs_ctx = 0;
  for (size_t s = 0; s < order.size(); ++s) {
    const fsweep_op_t& o = ops[order[s]];
    p->leaf.push_back(o);
    OpK& k = P.ops[s];
    k.kind = o.kind;
    k.n_out = o.n_out;
    k.n_in = o.n_in;
    k.K = o.n_sections;
    k.flags = o.flags;
    k.row_len = row_len_of(o);
    k.h_off = h_total;
    h_total += kind_is_diag(o.kind) ? 1 : o.n_in;
    if (o.kind == FSWEEP_OP_SOS || o.kind == FSWEEP_OP_PSOS) P.needs_ctx = 1;
    k.acc_off = acc_total;
    k.row_off = 0;
    k.def_off = -1;
    k.acc_mode = ACC_NONE;
    if (o.flags & FSWEEP_F_GRAD) {
      if (kind_is_table(o.kind)) {
        k.acc_mode = ACC_TABLE;
      } else {
        k.acc_mode = ACC_SMEM;
        acc_total += o.n_out * k.row_len;
      }
    }
  }
  // ---- wide flagship shape -> CTA-per-bin kernels: [GAIN N x 1] RECURSION(diagonal, no gradients ; GAIN N x N) [GAIN 1 x N]
  {
    const char* no_cta = getenv("FSWEEP_DISABLE_CTA");
    bool ok = !p->items && (dtype == FSWEEP_C128 || !(no_cta && no_cta[0] == '1')) && width > 32 && width <= 64 && rec >= 0 &&
              pre.size() == 1 && post.size() == 1 && fb.size() == 1 && ops[fb[0]].kind == FSWEEP_OP_GAIN &&
              ops[pre[0]].kind == FSWEEP_OP_GAIN && ops[post[0]].kind == FSWEEP_OP_GAIN && ops[pre[0]].n_in == 1 &&
              ops[post[0]].n_out == 1 && ops[pre[0]].n_out == rec_n && ops[post[0]].n_in == rec_n && rec_in == rec_n &&
              ff.size() <= 8;
    for (int i : ff) ok = ok && kind_is_diag(ops[i].kind) && !(ops[i].flags & FSWEEP_F_GRAD);
    if (ok) {
      p->cta = true;
      const char* tc_env = getenv("FSWEEP_CTA_TC");
      p->cta_tc = dtype == FSWEEP_C64 && (tc_env ? tc_env[0] == '1' : FSWEEP_CTA_TC_DEFAULT);
      memset(&p->loop, 0, sizeof(p->loop));
      p->loop.pre = slot_of[pre[0]];
      p->loop.ff_begin = slot_of[ff[0]];
      p->loop.n_ff = (int)ff.size();
      p->loop.fb = slot_of[fb[0]];
      p->loop.post = slot_of[post[0]];
    }
  }
  if (wide64 && !p->cta) {
    delete p;
    return fail(FSWEEP_E_UNSUPPORTED,
                "channel width %d > 32 in float64 is supported for the FDN shape only ([Gain N x 1] Recursion(diagonal "
                "chain without gradients ; N x N matrix) [Gain 1 x N]): complex128 rows exceed the register file of the "
                "row-distributed kernels", width);
  }
  // shared-memory accumulator budget: spill the largest rows to global atomics until it fits (the CTA kernels keep
  // every accumulator in registers: nothing to spill)
  for (;;) {
    acc_per_lane = 0;
    int big = -1;
    for (int s = 0; s < P.n_ops; ++s)
      if (P.ops[s].acc_mode == ACC_SMEM) {
        P.ops[s].row_off = acc_per_lane;
        acc_per_lane += P.ops[s].row_len;
        if (big < 0 || P.ops[s].row_len > P.ops[big].row_len) big = s;
      }
    if (p->cta || (size_t)acc_per_lane * BLOCK * real_sz <= SMEM_ACC_BUDGET || big < 0) break;
    P.ops[big].acc_mode = ACC_GLOBAL;
  }
  // section cascades too large for shared memory: defer their coefficient gradient to fsweep_sos_defer_kernel
  // (FSWEEP_DISABLE_DEFER=1, tests: keep the in-kernel global atomics)
  P.def_stride = 0;
  {
    const char* no_defer = getenv("FSWEEP_DISABLE_DEFER");
    if (!(no_defer && no_defer[0] == '1') && !p->items)
      for (int s = 0; s < P.n_ops; ++s)
        if (P.ops[s].acc_mode == ACC_GLOBAL && (P.ops[s].kind == FSWEEP_OP_SOS || P.ops[s].kind == FSWEEP_OP_PSOS) &&
            P.ops[s].K <= DEF_MAX_K) {
          P.ops[s].def_off = P.def_stride;
          P.def_stride += P.ops[s].n_in + P.ops[s].n_out;
        }
  }
  for (int s = 0; s < P.n_ops; ++s) {
    if (P.ops[s].acc_mode == ACC_GLOBAL) p->any_global = true;
    if (P.ops[s].acc_mode == ACC_SMEM || P.ops[s].acc_mode == ACC_GLOBAL) p->any_acc = true;
  }
  P.acc_per_lane = acc_per_lane;
  P.acc_total = acc_total;
  P.h_total = h_total;

  const std::vector<int>& first_chain = !pre.empty() ? pre : (rec >= 0 ? ff : post);
  const std::vector<int>& last_chain = !post.empty() ? post : (rec >= 0 ? ff : pre);
  P.in_ch = ops[first_chain.front()].n_in;
  P.out_ch = ops[last_chain.back()].n_out;

  // ---- step tables
  auto S = [&](int flat, unsigned fl) {
    Step s;
    s.op = (unsigned char)slot_of[flat];
    s.flags = (unsigned char)fl;
    return s;
  };
  int nf = 0, nm = 0, nb = 0, nr = 0, saves = 0;
  for (int i : pre) P.fsteps[nf++] = S(i, 0);
  for (size_t j = 0; j < ff.size(); ++j) P.fsteps[nf++] = S(ff[j], j + 1 == ff.size() ? ST_SOLVE : 0);
  for (int i : post) P.fsteps[nf++] = S(i, 0);
  for (size_t j = 0; j < fb.size(); ++j)
    P.msteps[nm++] = S(fb[j], (j == 0 && !kind_is_diag(ops[fb[j]].kind)) ? ST_IDENT : 0);
  for (int i : ff) P.msteps[nm++] = S(i, 0);

  for (int i : pre) P.bsteps[nb++] = S(i, ST_SAVE), ++saves;
  for (size_t j = 0; j < ff.size(); ++j)
    P.bsteps[nb++] = S(ff[j], (j == 0 ? ST_SAVE_X : 0) | (j + 1 == ff.size() ? ST_SOLVE : 0));
  for (size_t j = 0; j < fb.size(); ++j) P.bsteps[nb++] = S(fb[j], ST_SAVE | (j + 1 == fb.size() ? ST_ADD_X : 0)), ++saves;
  for (int i : ff) P.bsteps[nb++] = S(i, ST_SAVE), ++saves;
  for (int i : post) P.bsteps[nb++] = S(i, ST_SAVE), ++saves;

  for (int j = (int)post.size() - 1; j >= 0; --j) P.rsteps[nr++] = S(post[j], RS_NEED_GIN);
  for (int j = (int)ff.size() - 1; j >= 0; --j)
    P.rsteps[nr++] = S(ff[j], RS_NEED_GIN | (j + 1 == (int)ff.size() ? RS_ADJ : 0) | (j == 0 ? RS_SAVE_G : 0));
  for (int j = (int)fb.size() - 1; j >= 0; --j) P.rsteps[nr++] = S(fb[j], (j > 0 ? RS_NEED_GIN : 0) | (j == 0 ? RS_RESTORE_G : 0));
  p->first_pre_rstep = -1;
  for (int j = (int)pre.size() - 1; j >= 0; --j) {
    if (j == 0) p->first_pre_rstep = nr;
    P.rsteps[nr++] = S(pre[j], j > 0 ? RS_NEED_GIN : 0);
  }
  P.n_fsteps = nf;
  P.n_msteps = nm;
  P.n_bsteps = nb;
  P.n_rsteps = nr;
  P.n_slots = saves + 1;

  // ---- FDN-loop pattern: [GAIN] RECURSION(ff: diagonal ops, fb: one GAIN) [GAIN]
  const char* no_loop = getenv("FSWEEP_DISABLE_LOOP_KERNEL");  // tests: force the generic interpreter
  if (!(no_loop && no_loop[0] == '1') && !p->items && rec >= 0 && pre.size() <= 1 && post.size() <= 1 && fb.size() == 1 &&
      ops[fb[0]].kind == FSWEEP_OP_GAIN &&
      (pre.empty() || ops[pre[0]].kind == FSWEEP_OP_GAIN) && (post.empty() || ops[post[0]].kind == FSWEEP_OP_GAIN) &&
      G <= 32 && (dtype == FSWEEP_C64 || G <= 16)) {
    bool ok = true;
    for (int i : ff) ok = ok && kind_is_diag(ops[i].kind);
    for (int s2 = 0; s2 < P.n_ops; ++s2) ok = ok && P.ops[s2].acc_mode != ACC_GLOBAL;
    if (ok) {
      p->loop_fast = true;
      memset(&p->loop, 0, sizeof(p->loop));
      p->loop.pre = pre.empty() ? -1 : slot_of[pre[0]];
      p->loop.ff_begin = slot_of[ff[0]];
      p->loop.n_ff = (int)ff.size();
      p->loop.fb = slot_of[fb[0]];
      p->loop.post = post.empty() ? -1 : slot_of[post[0]];
      // thread-per-bin variant: float32, every width <= 8, few accumulators, gradients only for a single
      // PGAIN / PDELAY in the diagonal chain
      const char* no_tpb = getenv("FSWEEP_DISABLE_TPB");
      bool tpb = !(no_tpb && no_tpb[0] == '1') && dtype == FSWEEP_C64 && width <= 8 && acc_total <= 256;
      for (int i : ff) {
        const bool wants = (ops[i].flags & FSWEEP_F_GRAD) != 0;
        const bool simple = ops[i].kind == FSWEEP_OP_PGAIN || ops[i].kind == FSWEEP_OP_PDELAY;
        if (wants && !(simple && ff.size() == 1)) tpb = false;
      }
      if (tpb) p->tpb_np = width <= 4 ? 4 : 8;
      const char* force = getenv("FSWEEP_FORCE_TPB");
      p->tpb_force = force && force[0] == '1';
      // compact thread-per-bin variant: additionally one input and one output channel through N x 1 / 1 x N gains,
      // and every accumulator must fit the register-resident set the kernel stages (NP*NP + 3*NP slots)
      const char* no_tpc = getenv("FSWEEP_DISABLE_TPC");
      const int np = width <= 4 ? 4 : 8;
      if (tpb && !(no_tpc && no_tpc[0] == '1') && pre.size() == 1 && post.size() == 1 && ops[pre[0]].n_in == 1 &&
          ops[post[0]].n_out == 1 && ops[pre[0]].n_out == rec_n && ops[post[0]].n_in == rec_n && rec_in == rec_n &&
          acc_total <= np * np + 3 * np)
        p->tpc_np = np;
      const char* forcec = getenv("FSWEEP_FORCE_TPC");
      p->tpc_force = forcec && forcec[0] == '1';
    }
  }

  // ---- TABLE-heavy program without recursion -> streaming kernels
  {
    const char* no_stream = getenv("FSWEEP_DISABLE_STREAM");
    bool ok = !(no_stream && no_stream[0] == '1') && !p->items && dtype == FSWEEP_C64 && rec < 0 && width <= SW;
    bool any_table = false;
    for (int s2 = 0; s2 < P.n_ops && ok; ++s2) {
      const OpK& o = P.ops[s2];
      const bool tabk = o.kind == FSWEEP_OP_TABLE || o.kind == FSWEEP_OP_PTABLE;
      any_table |= tabk;
      ok = tabk || o.kind == FSWEEP_OP_PGAIN || (o.kind == FSWEEP_OP_GAIN && o.acc_mode == ACC_NONE);
      ok = ok && o.acc_mode != ACC_GLOBAL;
    }
    if (ok && any_table) {
      StreamInfo& S = p->sinfo;
      memset(&S, 0, sizeof(S));
      S.n_ops = P.n_ops;
      int off = 0, st = 0, pg = 0;
      for (int s2 = 0; s2 < P.n_ops; ++s2) {
        const OpK& o = P.ops[s2];
        S.tab_off[s2] = -1;
        S.pg_off[s2] = -1;
        S.row_bytes[s2] = 0;
        if (o.kind == FSWEEP_OP_TABLE || o.kind == FSWEEP_OP_PTABLE) {
          S.row_bytes[s2] = (o.kind == FSWEEP_OP_TABLE ? o.n_out * o.n_in : o.n_out) * 8;
          S.tab_off[s2] = off;  // in units of ONE bin; scaled by tb at launch
          off += S.row_bytes[s2];
        }
        if (o.kind == FSWEEP_OP_PGAIN && o.acc_mode == ACC_SMEM) {
          S.pg_off[s2] = pg;
          pg += o.n_out;
        }
        S.st_off[s2] = st;
        st += o.n_in;
      }
      S.st_off[P.n_ops] = st;
      S.st_total = st + P.out_ch;
      S.bytes_per_bin = off;
      S.n_pgain_acc = pg;
      p->stream = S.st_total <= S_MAXST;
    }
  }

  *out = p;
  return FSWEEP_OK;
}

extern "C" int fsweep_plan_destroy(fsweep_plan_t* plan) {
  delete plan;
  return FSWEEP_OK;
}

extern "C" int fsweep_plan_num_coeffs(const fsweep_plan_t* plan) { return plan ? plan->n_coeffs : 0; }

extern "C" int64_t fsweep_plan_coeff_numel(const fsweep_plan_t* plan, int slot, int64_t M) {
  if (!plan || slot < 0 || slot >= plan->n_coeffs) return -1;
  const fsweep_op_t& o = plan->leaf[slot];
  switch (o.kind) {
    case FSWEEP_OP_GAIN:
    case FSWEEP_OP_DELAY:
      return (int64_t)o.n_out * o.n_in;
    case FSWEEP_OP_PGAIN:
    case FSWEEP_OP_PDELAY:
      return o.n_out;
    case FSWEEP_OP_SOS:
      return (int64_t)o.n_sections * o.n_in * o.n_out * 16;
    case FSWEEP_OP_PSOS:
      return (int64_t)o.n_sections * o.n_out * 16;
    case FSWEEP_OP_TABLE:
      return M * o.n_out * o.n_in;
    case FSWEEP_OP_PTABLE:
      return M * o.n_out;
    default:
      return -1;
  }
}

namespace {

// One thread per bin only pays when there are enough bins to fill the machine with warps (a bin is one THREAD
// there, a quarter-warp or more in the row-distributed kernels).  Measured on B200, 8x8 FDN, 48001 bins:
// forward 20 us (thread-per-bin) vs 29 us (row-distributed); backward 69 us vs 49 us (profiles/r01_notes.md).
constexpr int64_t TPB_MIN_BINS_FWD = 32768;
constexpr int64_t TPB_MIN_BINS_BWD = 262144;
bool use_tpb(const fsweep_plan* p, int64_t n_bins, bool bwd) {
  if (!p->tpb_np) return false;
  return p->tpb_force || n_bins >= (bwd ? TPB_MIN_BINS_BWD : TPB_MIN_BINS_FWD);
}

// The compact kernels are the better choice as soon as there is about one bin per resident thread.
constexpr int64_t TPC_MIN_BINS = 16384;
bool use_tpc(const fsweep_plan* p, int64_t n_bins) {
  if (!p->tpc_np || p->tpb_force) return false;
  return p->tpc_force || n_bins >= TPC_MIN_BINS;
}
// Register-matrix variant (fsweep_tpr.cuh) of the compact kernels: the default for loop widths 5..8; FSWEEP_TPC_V1=1 (tests,
// A/B measurements) keeps the shared-memory-matrix kernels of fsweep_tpc.cuh.  Read per call.
bool use_tpr(const fsweep_plan* p) {
  if (p->tpc_np != 8) return false;
  const char* v1 = getenv("FSWEEP_TPC_V1");
  return !(v1 && v1[0] == '1');
}
// one block of TPR_BLOCK threads per SM, all bins in one wave when they fit
int tpr_grid(fsweep_plan* p, int64_t n_bins, cudaError_t* err) {
  *err = cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    if (p->num_sms == 0) {
      int dev = 0;
      if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return 0;
      if ((*err = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return 0;
    }
  }
  return (int)std::min<int64_t>((n_bins + TPR_BLOCK - 1) / TPR_BLOCK, p->num_sms);
}
int tpc_grid(int64_t n_bins) { return (int)std::min<int64_t>((n_bins + TPC_BLOCK - 1) / TPC_BLOCK, MAX_GRID); }

int cc_of(int64_t ncols) { return ncols == 1 ? 1 : 4; }

size_t smem_fwd(const fsweep_plan* p) {
  return (size_t)p->prog.h_total * BLOCK * 2 * (p->dtype == FSWEEP_C64 ? 4 : 8) + (size_t)p->prog.n_ops * BLOCK * 4;
}
size_t smem_bwd(const fsweep_plan* p, int cc) {
  const size_t rs = p->dtype == FSWEEP_C64 ? 4 : 8;
  return smem_fwd(p) + (size_t)p->prog.n_slots * cc * BLOCK * 2 * rs + (size_t)p->prog.acc_per_lane * BLOCK * rs;
}

int grid_cap(int64_t n_bins, int G) {
  const int64_t per_block = BLOCK / G;
  int64_t need = (n_bins + per_block - 1) / per_block;
  return (int)std::min<int64_t>(need, MAX_GRID);
}

int cta_grid(fsweep_plan* p, bool bwd, int64_t n_bins, cudaError_t* err) {
  *err = cudaSuccess;
  std::lock_guard<std::mutex> lk(p->mu);
  if (p->num_sms == 0) {
    int dev = 0;
    if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return 0;
    if ((*err = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return 0;
  }
  if (p->cta_blocks_per_sm[bwd] == 0) {
    int n = 0;
    if ((*err = occupancy_cta(p->dtype, bwd, p->cta_tc, &n)) != cudaSuccess) return 0;
    if (n < 1) {
      *err = cudaErrorLaunchOutOfResources;
      return 0;
    }
    // The occupancy calculator answers 1 for a kernel that allocates tensor memory, but blocks do share an SM as long
    // as their TMEM columns fit (64 of 512 each here): registers and shared memory allow 5 (__launch_bounds__).  A
    // block that finds no room simply starts later; the bin loop is grid-strided either way.
    if (p->cta_tc) n = std::max(n, tc::BLOCKS_PER_SM);
    p->cta_blocks_per_sm[bwd] = n;
    if (getenv("FSWEEP_DEBUG")) fprintf(stderr, "[fsweep] cta kernel bwd=%d tc=%d: %d blocks per SM, %d SMs\n", (int)bwd, (int)p->cta_tc, n, p->num_sms);
  }
  const int64_t resident = (int64_t)p->cta_blocks_per_sm[bwd] * p->num_sms;
  return (int)std::min<int64_t>(std::min<int64_t>(resident, n_bins), grid_cap(n_bins, p->G));
}

// per-call geometry of the streaming kernels; returns false when this call cannot use them
// TMA (bulk-copy) variant of the streaming kernels: every table block (and table-gradient block) of the call must start
// on a 16-byte boundary — true for whole tables, not for every bin shard.  OPT-IN (FSWEEP_STREAM_TMA=1): measured on a
// B200 it is 6 - 30 % SLOWER than the cp.async path (profiles/r02_notes.md) — the ring of three tile buffers costs
// shared memory, shared memory per bin is what bounds the bins in flight per SM, and the kernel is bound by the
// LDS -> FMA latency of the bins in flight, not by the DRAM latency the ring hides.
bool stream_tma_ok(const ProgK& P, const StreamInfo& S, int64_t bin_begin, bool bwd) {
  const char* env = getenv("FSWEEP_STREAM_TMA");  // read per call: tests toggle it
  if (!(env && env[0] == '1')) return false;
  for (int i = 0; i < S.n_ops; ++i) {
    if (S.row_bytes[i] <= 0 || P.ops[i].coef == nullptr) continue;
    if ((reinterpret_cast<uintptr_t>(P.ops[i].coef) + (uintptr_t)bin_begin * S.row_bytes[i]) % 16) return false;
    if (bwd && P.ops[i].acc_mode == ACC_TABLE &&
        (reinterpret_cast<uintptr_t>(P.ops[i].gtab) + (uintptr_t)bin_begin * S.row_bytes[i]) % 16)
      return false;
  }
  return true;
}

bool stream_setup(const fsweep_plan* p, int64_t q, bool bwd, bool tma, StreamInfo* S, size_t* smem) {
  if (!p->stream || q < 1 || q > 16 || (q & (q - 1)) != 0) return false;
  *S = p->sinfo;
  S->qc = (int)q;
  // one thread per (bin, column).  cp.async path: small blocks (shared memory per bin bounds the warps per SM); TMA
  // path: tiles of >= 16 bins so that one bulk copy moves a few KB
  static const int tma_threads = [] {
    const char* e = getenv("FSWEEP_STREAM_TMA_THREADS");
    const int v = e ? atoi(e) : 64;
    return (v == 32 || v == 64 || v == 128) ? v : 64;
  }();
  S->threads = tma ? std::max(tma_threads, 2 * (int)q) : 32;
  S->tb = S->threads / (int)q;
  if (tma && (S->tb & 1)) return false;  // block sizes must be multiples of 16 bytes (rows are multiples of 8)
  for (int i = 0; i < S->n_ops; ++i)
    if (S->tab_off[i] >= 0) S->tab_off[i] *= S->tb;  // blocks of tb rows per op inside a stage
  const size_t stage = (size_t)S->tb * S->bytes_per_bin;
  const size_t n_state = bwd ? (size_t)S->st_total : 2 * SW;
  *smem = (tma ? S_TMA_STAGES : S_STAGES) * stage + n_state * S->threads * 8 +
          (bwd ? (size_t)2 * SW * S->threads * 8 + (size_t)S->n_pgain_acc * 4 : 0) + 16 +
          (tma ? 16 + 8 * S_TMA_STAGES : 0);
  return *smem <= 200 * 1024;
}

// FSWEEP_STREAM_FWD selects the forward streaming kernel (read per call: tests toggle it).  Measured on a B200 on the
// FIR chain of tools/measure_table_sweep.py (48 001 bins, 936 B of tables per bin, 4 columns; profiles/r02_notes.md):
//   "reg"  (default) thread per (bin, column), state in registers (fsweep_streamr.cuh), cp.async tiles: 28.7 us
//          ... with FSWEEP_STREAM_REG_TMA=1 the tiles come through the bulk-copy engine, two buffers: 32.8 us
//   "smem" thread per (bin, column), state in shared memory (fsweep_stream.cuh): 34.8 us
//   "warp" one warp per bin, state distributed over the lanes (fsweep_streamw.cuh): 49.1 us
int stream_fwd_kind() {
  const char* env = getenv("FSWEEP_STREAM_FWD");
  if (env && env[0] == 'w') return 2;
  if (env && env[0] == 's') return 0;
  return 1;
}

// geometry of a call of the forward streaming kernels fsweep_streamr.cuh / fsweep_streamw.cuh
bool streamr_setup(const fsweep_plan* p, const ProgK& P, int64_t q, int64_t bin_begin, bool warp_per_bin, StreamInfo* S,
                   StreamRInfo* R, size_t* smem, int* w, bool* tma) {
  if (!p->stream || q < 1 || q > 16 || (q & (q - 1)) != 0) return false;
  *S = p->sinfo;
  S->qc = (int)q;
  static const int r_threads = [] {
    const char* e = getenv("FSWEEP_STREAMR_THREADS");
    const int v = e ? atoi(e) : 64;
    return (v == 64 || v == 128 || v == 256) ? v : 64;
  }();
  S->threads = warp_per_bin ? SWARP_THREADS : r_threads;
  S->tb = warp_per_bin ? 16 : S->threads / (int)q;
  int width = std::max(P.in_ch, P.out_ch), units = 0;
  // bulk copies need 16-byte aligned sources and sizes: an even number of bins per tile makes every size a multiple
  // of 16 (rows are multiples of 8); the sources are aligned for whole tables, not for every bin shard
  const char* want_tma = getenv("FSWEEP_STREAM_REG_TMA");
  *tma = (S->tb % 2 == 0);  // (so far: "aligned")
  for (int i = 0; i < S->n_ops; ++i) {
    width = std::max(width, std::max(P.ops[i].n_in, P.ops[i].n_out));
    R->pad_units[i] = 0;
    R->pad_off[i] = 0;
    if (S->tab_off[i] < 0) continue;
    R->pad_units[i] = S->row_bytes[i] / 8;
    R->pad_off[i] = units;
    units += S->tb * R->pad_units[i];
    if ((reinterpret_cast<uintptr_t>(P.ops[i].coef) + (uintptr_t)bin_begin * S->row_bytes[i]) % 16) *tma = false;
  }
  R->tile_units = units;
  R->aligned16 = *tma ? 1 : 0;
  *tma = *tma && (warp_per_bin || (want_tma && want_tma[0] == '1'));
  *smem = (size_t)units * 8 * (*tma ? 2 : 1) + 16;
  *w = width <= 4 ? 4 : (width <= 8 ? 8 : 16);
  return width <= 16 && *smem <= 200 * 1024;
}

template <typename F>
cudaError_t by_group(int G, F&& f) {
  switch (G) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 4: return f(std::integral_constant<int, 4>());
    case 8: return f(std::integral_constant<int, 8>());
    case 16: return f(std::integral_constant<int, 16>());
    case 32: return f(std::integral_constant<int, 32>());
    default: return f(std::integral_constant<int, 64>());
  }
}

// persistent grid: resident blocks per SM (occupancy query, cached) x SM count, capped by the work
int pick_grid(fsweep_plan* p, int cc, bool bwd, size_t smem, int64_t n_bins, cudaError_t* err, bool loop = false) {
  const int ci = loop ? 2 : (cc == 1 ? 0 : 1);
  *err = cudaSuccess;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    if (p->num_sms == 0) {
      int dev = 0;
      if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return 0;
      if ((*err = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return 0;
    }
    if (p->blocks_per_sm[ci][bwd] == 0) {
      int n = 0;
      const int dtype = p->dtype;
      if (loop)
        *err = by_group(p->G, [&](auto g) { return occupancy_loop<decltype(g)::value>(dtype, bwd, smem, &n); });
      else
        *err = by_group(p->G, [&](auto g) { return occupancy<decltype(g)::value>(dtype, cc, bwd, smem, &n); });
      if (*err != cudaSuccess) return 0;
      if (n < 1) {
        *err = cudaErrorLaunchOutOfResources;
        return 0;
      }
      p->blocks_per_sm[ci][bwd] = n;
    }
  }
  int64_t resident = (int64_t)p->blocks_per_sm[ci][bwd] * p->num_sms;
  return (int)std::min<int64_t>(resident, grid_cap(n_bins, p->G));
}

size_t slot_bytes(const fsweep_plan* plan, int slot) {
  const int kind = plan->leaf[slot].kind;
  const size_t rs = plan->dtype == FSWEEP_C64 ? 4 : 8;
  const size_t n = (size_t)fsweep_plan_coeff_numel(plan, slot, plan->prog.nfft / 2 + 1);
  if (kind == FSWEEP_OP_DELAY || kind == FSWEEP_OP_PDELAY) return n * 8;
  return n * (kind_is_table(kind) ? 2 * rs : rs);
}

// Per-item launches (fsweep_op_t::per_item): `batch` is the item count, the grid gets one y-slice per item and the
// x-extent is what keeps items * grid.x near the resident block count.  Fills the item strides of P.
int items_setup(fsweep_plan* plan, ProgK& P, int64_t batch, int* grid_x) {
  if (batch > 65535 || batch > MAX_GRID) return fail(FSWEEP_E_UNSUPPORTED, "per-item coefficients: at most %d items per call", MAX_GRID);
  for (int s = 0; s < plan->n_coeffs; ++s) {
    const long long b = plan->leaf[s].per_item ? (long long)slot_bytes(plan, s) : 0;
    P.ops[s].coef_is = b;
    P.ops[s].gtab_is = b;
  }
  const int gx = std::max<int>(1, std::min<int64_t>(MAX_GRID / batch, (*grid_x + batch - 1) / batch));
  *grid_x = std::min(*grid_x, gx);
  return FSWEEP_OK;
}

int check_common(const fsweep_plan* plan, const void* const* coeffs, const void* x, int64_t batch, int64_t cols,
                 int64_t bin_begin, int64_t n_bins, int epilogue) {
  if (!plan || !coeffs || !x) return fail(FSWEEP_E_BADARG, "null plan / coeffs / x");
  if (batch < 1 || cols < 1 || n_bins < 0 || bin_begin < 0) return fail(FSWEEP_E_BADARG, "bad batch/cols/bins");
  if (batch * cols > (1 << 20)) return fail(FSWEEP_E_BADARG, "batch*cols too large");
  if (bin_begin + n_bins > plan->prog.nfft / 2 + 1)
    return fail(FSWEEP_E_BADARG, "bin range [%lld, %lld) exceeds nfft/2+1", (long long)bin_begin, (long long)(bin_begin + n_bins));
  if (epilogue != FSWEEP_EPI_NONE && epilogue != FSWEEP_EPI_ABS) return fail(FSWEEP_E_BADARG, "bad epilogue %d", epilogue);
  for (int s = 0; s < plan->n_coeffs; ++s)
    if (!coeffs[s]) return fail(FSWEEP_E_BADARG, "coefficient slot %d is null", s);
  return FSWEEP_OK;
}

}  // namespace

extern "C" const char* fsweep_plan_kernel_family(const fsweep_plan_t* plan, int64_t n_bins, int backward) {
  if (!plan) return "";
  if (plan->cta && plan->cta_tc) return backward ? "fsweep_cta_kernel<bwd,tc> (tcgen05 LU)" : "fsweep_cta_kernel<fwd,tc> (tcgen05 LU)";
  if (plan->cta) return backward ? "fsweep_cta_kernel<bwd>" : "fsweep_cta_kernel<fwd>";
  if (plan->stream) return backward ? "fsweep_stream_kernel<bwd> (batch*cols a power of two <= 16)" : "fsweep_streamr_kernel (fwd, state in registers; batch*cols a power of two <= 16)";
  if (use_tpc(plan, n_bins) && use_tpr(plan)) return backward ? "fsweep_tpr_kernel<bwd> (tpc family, matrix in registers)" : "fsweep_tpr_kernel<fwd> (tpc family, matrix in registers)";
  if (use_tpc(plan, n_bins)) return backward ? "fsweep_tpc_kernel<NP,bwd>" : "fsweep_tpc_kernel<NP,fwd>";
  if (use_tpb(plan, n_bins, backward != 0)) return backward ? "fsweep_tpb_bwd_kernel" : "fsweep_tpb_fwd_kernel";
  if (plan->loop_fast) return backward ? "fsweep_loop_bwd_kernel" : "fsweep_loop_fwd_kernel";
  return backward ? "fsweep_bwd_kernel" : "fsweep_fwd_kernel";
}

namespace {
// [per-block partial sums][flat global accumulator][pad][loss partials] — then the deferral buffer
size_t ws_base_bytes(const fsweep_plan* plan, int64_t n_bins, int64_t items = 1) {
  const size_t rs = plan->dtype == FSWEEP_C64 ? 4 : 8;
  const size_t grid = (size_t)grid_cap(n_bins, plan->G);
  size_t partial = grid * (size_t)items * (size_t)plan->prog.acc_per_lane * plan->G * rs;  // (per-item launches: [item][block])
  size_t gacc = (size_t)items * (size_t)plan->prog.acc_total * rs;
  return ((partial + 255) / 256) * 256 + ((gacc + 255) / 256) * 256 + 256 + LOSS_PARTIAL_BYTES;
}
// deferral records, then one response table per deferred op
size_t ws_defer_bytes(const fsweep_plan* plan, int64_t batch, int64_t cols, int64_t n_bins) {
  const size_t rs = plan->dtype == FSWEEP_C64 ? 4 : 8;
  size_t b = (((size_t)plan->prog.def_stride * (size_t)(batch * cols) * (size_t)n_bins * 2 * rs + 255) / 256) * 256;
  for (int s = 0; s < plan->prog.n_ops; ++s) {
    const OpK& o = plan->prog.ops[s];
    if (o.def_off < 0) continue;
    const size_t row = o.kind == FSWEEP_OP_PSOS ? (size_t)o.n_out : (size_t)o.n_out * o.n_in;
    b += ((row * (size_t)n_bins * 2 * rs + 255) / 256) * 256;
  }
  return b;
}
}  // namespace

extern "C" size_t fsweep_workspace_bytes(const fsweep_plan_t* plan, int64_t batch, int64_t cols, int64_t n_bins) {
  if (!plan) return 0;
  return ws_base_bytes(plan, n_bins, plan->items ? batch : 1) + ws_defer_bytes(plan, batch, cols, n_bins);
}

namespace {

// validates a fused criterion and fills the kernel-side fields; returns the internal epilogue code in *epi
int setup_criterion(const fsweep_plan* plan, const fsweep_criterion_t* crit, int64_t n_bins, int64_t items, void* workspace,
                    size_t workspace_bytes, SweepArgs& A, int* epi) {
  if (!crit) return fail(FSWEEP_E_BADARG, "null criterion");
  if (crit->kind != FSWEEP_CRIT_MSE && crit->kind != FSWEEP_CRIT_MSE_CHSUM)
    return fail(FSWEEP_E_BADARG, "bad criterion kind %d", crit->kind);
  if (!crit->target || !crit->loss) return fail(FSWEEP_E_BADARG, "criterion: null target / loss");
  const size_t need = ws_base_bytes(plan, n_bins, items);
  if (!workspace || workspace_bytes < need)
    return fail(FSWEEP_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  *epi = crit->kind == FSWEEP_CRIT_MSE ? EPI_ABS_MSE : EPI_ABSSUM_MSE;
  A.tgt = crit->target;
  A.tbs = crit->target_batch_stride;
  A.crit_scale = crit->scale;
  A.loss_partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + (need - LOSS_PARTIAL_BYTES));
  return FSWEEP_OK;
}

void launch_loss_finalize(int dtype, int n_blocks, int n_items, const SweepArgs& A, void* loss, cudaStream_t st) {
  FinalizeArgs F;
  memset(&F, 0, sizeof(F));
  F.n_ops = 0;
  F.n_blocks = n_blocks;
  F.n_items = n_items;
  F.loss_partial = A.loss_partial;
  F.loss = loss;
  F.crit_scale = A.crit_scale;
  if (dtype == FSWEEP_C64)
    fsweep_finalize_kernel<float><<<dim3(1, 1), 128, 0, st>>>(F);
  else
    fsweep_finalize_kernel<double><<<dim3(1, 1), 128, 0, st>>>(F);
}

int forward_impl(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x, int64_t x_batch_stride, void* y,
                 int64_t y_batch_stride, int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins, int epilogue,
                 const fsweep_criterion_t* crit, void* workspace, size_t workspace_bytes, void* stream);

}  // namespace

extern "C" int fsweep_forward(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x,
                              int64_t x_batch_stride, void* y, int64_t y_batch_stride, int64_t batch, int64_t cols,
                              int64_t bin_begin, int64_t n_bins, int epilogue, void* stream) {
  if (epilogue != FSWEEP_EPI_NONE && epilogue != FSWEEP_EPI_ABS) return fail(FSWEEP_E_BADARG, "bad epilogue %d", epilogue);
  return forward_impl(plan_c, coeffs, x, x_batch_stride, y, y_batch_stride, batch, cols, bin_begin, n_bins, epilogue,
                      nullptr, nullptr, 0, stream);
}

extern "C" int fsweep_forward_loss(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x,
                                   int64_t x_batch_stride, const fsweep_criterion_t* crit, int64_t batch,
                                   int64_t bin_begin, int64_t n_bins, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (!crit) return fail(FSWEEP_E_BADARG, "null criterion");
  if (n_bins <= 0) return fail(FSWEEP_E_BADARG, "empty bin range in a fused criterion");
  return forward_impl(plan_c, coeffs, x, x_batch_stride, nullptr, 0, batch, 1, bin_begin, n_bins, FSWEEP_EPI_ABS, crit,
                      workspace, workspace_bytes, stream);
}

namespace {
int forward_impl(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x, int64_t x_batch_stride, void* y,
                 int64_t y_batch_stride, int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins, int epilogue,
                 const fsweep_criterion_t* crit, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  fsweep_plan* plan = const_cast<fsweep_plan*>(plan_c);
  int r = check_common(plan, coeffs, x, batch, cols, bin_begin, n_bins, epilogue);
  if (r) return r;
  if (!y && !crit) return fail(FSWEEP_E_BADARG, "null y");
  if (n_bins == 0) return FSWEEP_OK;
  ProgK P = plan->prog;
  for (int s = 0; s < plan->n_coeffs; ++s) P.ops[s].coef = coeffs[s];
  SweepArgs A;
  memset(&A, 0, sizeof(A));
  A.x = x;
  A.xbs = x_batch_stride;
  A.y = y;
  A.ybs = y_batch_stride;
  A.batch = (int)batch;
  A.cols = (int)cols;
  A.bin_begin = bin_begin;
  A.n_bins = n_bins;
  A.epilogue = epilogue;
  const int64_t items = plan->items ? batch : 1;
  if (crit && (r = setup_criterion(plan, crit, n_bins, items, workspace, workspace_bytes, A, &A.epilogue))) return r;
  const int cc = cc_of(plan->items ? cols : batch * cols);
  const bool loop = plan->loop_fast;
  LaunchCfg cfg;
  cfg.smem = plan->cta ? 0 : smem_fwd(plan);
  if (cfg.smem > 220 * 1024) return fail(FSWEEP_E_UNSUPPORTED, "forward needs %zu bytes of shared memory", cfg.smem);
  cfg.stream = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  cfg.grid = plan->cta ? 0 : pick_grid(plan, cc, false, cfg.smem, n_bins, &e, loop);
  if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "occupancy query: %s", cudaGetErrorString(e));
  if (plan->items && (r = items_setup(plan, P, batch, &cfg.grid))) return r;
  cfg.items = (int)items;
  const int dtype = plan->dtype;
  StreamInfo SI;
  StreamRInfo RI;
  size_t ssmem = 0;
  bool stream_tma = false;
  int rw = 0;
  bool rtma = false;
  if (plan->cta) {
    cfg.grid = cta_grid(plan, false, n_bins, &e);
    if (e == cudaSuccess) e = launch_cta(plan->dtype, false, plan->cta_tc, cfg.grid, cfg.stream, P, plan->loop, A, plan->G);
  } else if (!crit && stream_fwd_kind() != 0 &&
             streamr_setup(plan, P, batch * cols, bin_begin, stream_fwd_kind() == 2, &SI, &RI, &ssmem, &rw, &rtma)) {
    const bool wpb = stream_fwd_kind() == 2;
    const size_t key = ssmem * 4 + (wpb ? 2 : 0) + (rtma ? 1 : 0);
    int bps = plan->streamr_bps_smem == key ? plan->streamr_bps : 0;
    if (bps == 0) {
      e = wpb ? occupancy_streamw(SI.qc, rtma, ssmem, &bps) : occupancy_streamr(rw, rtma, SI.threads, ssmem, &bps);
      if (e == cudaSuccess) {
        plan->streamr_bps = bps;
        plan->streamr_bps_smem = key;
      }
    }
    if (e == cudaSuccess) {
      const int64_t tiles = (n_bins + SI.tb - 1) / SI.tb;
      cfg.grid = (int)std::min<int64_t>(tiles, (int64_t)std::max(1, bps) * std::max(1, plan->num_sms ? plan->num_sms : 148));
      e = wpb ? launch_streamw(SI.qc, rtma, cfg.grid, ssmem, cfg.stream, P, SI, RI, A)
              : launch_streamr(rw, rtma, cfg.grid, SI.threads, ssmem, cfg.stream, P, SI, RI, A);
    }
  } else if (!crit && ((stream_tma = plan->stream && stream_tma_ok(P, plan->sinfo, bin_begin, false) &&
                                     stream_setup(plan, batch * cols, false, true, &SI, &ssmem)) ||
                       stream_setup(plan, batch * cols, false, false, &SI, &ssmem))) {
    int bps = plan->stream_bps_smem[0] == ssmem ? plan->stream_bps[0] : 0;
    if (bps == 0) {
      e = occupancy_stream(false, stream_tma, SI.threads, ssmem, &bps);
      if (e == cudaSuccess) {
        plan->stream_bps[0] = bps;
        plan->stream_bps_smem[0] = ssmem;
      }
    }
    if (e == cudaSuccess) {
      const int64_t tiles = (n_bins + SI.tb - 1) / SI.tb;
      cfg.grid = (int)std::min<int64_t>(tiles, (int64_t)std::max(1, bps) * std::max(1, plan->num_sms ? plan->num_sms : 148));
      e = launch_stream(false, stream_tma, cfg.grid, ssmem, cfg.stream, P, SI, A, plan->G);
    }
  } else if (use_tpc(plan, n_bins) && use_tpr(plan)) {
    cfg.grid = tpr_grid(plan, n_bins, &e);
    if (e == cudaSuccess) e = launch_tpr(plan->tpc_np, false, cfg.grid, cfg.stream, P, plan->loop, A, plan->G);
  } else if (use_tpc(plan, n_bins)) {
    cfg.grid = tpc_grid(n_bins);
    e = launch_tpc(plan->tpc_np, false, cfg.grid, cfg.stream, P, plan->loop, A, plan->G);
  } else if (use_tpb(plan, n_bins, false)) {
    const LoopInfo L = plan->loop;
    cfg.grid = (int)std::min<int64_t>((n_bins + TPB_BLOCK - 1) / TPB_BLOCK, grid_cap(n_bins, plan->G));
    e = launch_tpb_fwd(plan->tpb_np, cfg.grid, cfg.stream, P, L, A);
  } else if (loop) {
    const LoopInfo L = plan->loop;
    e = by_group(plan->G, [&](auto g) { return launch_loop_fwd<decltype(g)::value>(dtype, cfg, P, L, A); });
  } else {
    e = by_group(plan->G, [&](auto g) { return launch_fwd<decltype(g)::value>(dtype, cc, cfg, P, A); });
  }
  if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "forward launch: %s", cudaGetErrorString(e));
  g_launches = 1;
  if (crit) {
    launch_loss_finalize(dtype, cfg.grid, (int)items, A, crit->loss, cfg.stream);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(FSWEEP_E_CUDA, "loss finalize launch: %s", cudaGetErrorString(e));
    g_launches = 2;
  }
  return FSWEEP_OK;
}
}  // namespace

namespace {
int backward_impl(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x, int64_t x_batch_stride,
                  const void* grad_y, int64_t gy_batch_stride, void* const* grad_coeffs, void* grad_x,
                  int64_t gx_batch_stride, int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins, int epilogue,
                  const fsweep_criterion_t* crit, void* workspace, size_t workspace_bytes, void* stream);
}

extern "C" int fsweep_backward(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x,
                               int64_t x_batch_stride, const void* grad_y, int64_t gy_batch_stride,
                               void* const* grad_coeffs, void* grad_x, int64_t gx_batch_stride, int64_t batch,
                               int64_t cols, int64_t bin_begin, int64_t n_bins, int epilogue, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (epilogue != FSWEEP_EPI_NONE && epilogue != FSWEEP_EPI_ABS) return fail(FSWEEP_E_BADARG, "bad epilogue %d", epilogue);
  return backward_impl(plan_c, coeffs, x, x_batch_stride, grad_y, gy_batch_stride, grad_coeffs, grad_x, gx_batch_stride,
                       batch, cols, bin_begin, n_bins, epilogue, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int fsweep_backward_loss(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x,
                                    int64_t x_batch_stride, const fsweep_criterion_t* crit, void* const* grad_coeffs,
                                    void* grad_x, int64_t gx_batch_stride, int64_t batch, int64_t bin_begin,
                                    int64_t n_bins, void* workspace, size_t workspace_bytes, void* stream) {
  if (!crit) return fail(FSWEEP_E_BADARG, "null criterion");
  return backward_impl(plan_c, coeffs, x, x_batch_stride, nullptr, 0, grad_coeffs, grad_x, gx_batch_stride, batch, 1,
                       bin_begin, n_bins, FSWEEP_EPI_ABS, crit, workspace, workspace_bytes, stream);
}

namespace {
int backward_impl(const fsweep_plan_t* plan_c, const void* const* coeffs, const void* x, int64_t x_batch_stride,
                  const void* grad_y, int64_t gy_batch_stride, void* const* grad_coeffs, void* grad_x,
                  int64_t gx_batch_stride, int64_t batch, int64_t cols, int64_t bin_begin, int64_t n_bins, int epilogue,
                  const fsweep_criterion_t* crit, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  fsweep_plan* plan = const_cast<fsweep_plan*>(plan_c);
  int r = check_common(plan, coeffs, x, batch, cols, bin_begin, n_bins, epilogue);
  if (r) return r;
  if (!grad_y && !crit) return fail(FSWEEP_E_BADARG, "null grad_y");
  if (n_bins == 0) return fail(FSWEEP_E_BADARG, "empty bin range in backward");
  bool any_deferred = false;
  for (int s = 0; s < plan->n_coeffs; ++s)
    if (plan->prog.ops[s].def_off >= 0 && grad_coeffs && grad_coeffs[s]) any_deferred = true;
  const int64_t items = plan->items ? batch : 1;
  const size_t need = ws_base_bytes(plan, n_bins, items) + (any_deferred ? ws_defer_bytes(plan, batch, cols, n_bins) : 0);
  if (need > 256 + LOSS_PARTIAL_BYTES && (!workspace || workspace_bytes < need))
    return fail(FSWEEP_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = plan->dtype == FSWEEP_C64 ? 4 : 8;

  ProgK P = plan->prog;
  bool any_acc_wanted = false;
  for (int s = 0; s < plan->n_coeffs; ++s) {
    P.ops[s].coef = coeffs[s];
    void* gp = grad_coeffs ? grad_coeffs[s] : nullptr;
    if (P.ops[s].acc_mode == ACC_TABLE) {
      P.ops[s].gtab = gp;
      if (!gp) P.ops[s].acc_mode = ACC_NONE;
    } else if (P.ops[s].acc_mode != ACC_NONE) {
      if (gp) any_acc_wanted = true;
      if (!gp) P.ops[s].def_off = -1;  // nobody wants this gradient: nothing to park either
    }
  }
  if (grad_x && plan->first_pre_rstep >= 0) P.rsteps[plan->first_pre_rstep].flags |= RS_NEED_GIN;

  const bool loop = plan->loop_fast;
  const int cc = loop ? 1 : cc_of(plan->items ? cols : batch * cols);
  LaunchCfg cfg;
  cfg.smem = plan->cta ? 0 : smem_bwd(plan, cc);
  if (cfg.smem > 220 * 1024) return fail(FSWEEP_E_UNSUPPORTED, "backward needs %zu bytes of shared memory", cfg.smem);
  cfg.stream = st;
  cudaError_t e = cudaSuccess;
  cfg.grid = plan->cta ? 0 : pick_grid(plan, cc, true, cfg.smem, n_bins, &e, loop);
  if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "occupancy query: %s", cudaGetErrorString(e));
  if (plan->items) {
    if ((r = items_setup(plan, P, batch, &cfg.grid))) return r;
    for (int s = 0; s < plan->n_coeffs; ++s)
      if (P.ops[s].acc_mode == ACC_TABLE && !plan->leaf[s].per_item)
        return fail(FSWEEP_E_UNSUPPORTED, "slot %d: the gradient of a TABLE shared by all items is not formed by a per-item launch", s);
  }
  cfg.items = (int)items;

  const size_t partial_bytes =
      (((size_t)grid_cap(n_bins, plan->G) * (size_t)items * P.acc_per_lane * plan->G * rs + 255) / 256) * 256;
  char* ws = reinterpret_cast<char*>(workspace);
  void* partial = ws;
  void* gacc = ws ? ws + partial_bytes : nullptr;

  SweepArgs A;
  memset(&A, 0, sizeof(A));
  A.x = x;
  A.xbs = x_batch_stride;
  A.gy = grad_y;
  A.gybs = gy_batch_stride;
  A.gx = grad_x;
  A.gxbs = gx_batch_stride;
  A.batch = (int)batch;
  A.cols = (int)cols;
  A.bin_begin = bin_begin;
  A.n_bins = n_bins;
  A.epilogue = epilogue;
  A.partial = partial;
  A.gacc = gacc;
  A.defer = any_deferred ? ws + ws_base_bytes(plan, n_bins) : nullptr;
  if (crit && (r = setup_criterion(plan, crit, n_bins, items, workspace, workspace_bytes, A, &A.epilogue))) return r;

  int launches = 0;
  if (plan->any_global) {
    e = cudaMemsetAsync(gacc, 0, (size_t)items * P.acc_total * rs, st);
    if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "memset: %s", cudaGetErrorString(e));
    ++launches;
  }
  // deferred section cascades: build their response tables first (the sweep kernel reads them as TABLE ops)
  ProgK Pdef = P;  // the program as the table / gradient kernels see it: original op kinds + table pointers
  if (any_deferred) {
    char* tabp = ws + ws_base_bytes(plan, n_bins) +
                 (((size_t)P.def_stride * (size_t)(batch * cols) * (size_t)n_bins * 2 * rs + 255) / 256) * 256;
    for (int s = 0; s < P.n_ops; ++s) {
      if (plan->prog.ops[s].def_off < 0) continue;
      const bool par = P.ops[s].kind == FSWEEP_OP_PSOS;
      const size_t row = par ? (size_t)P.ops[s].n_out : (size_t)P.ops[s].n_out * P.ops[s].n_in;
      const size_t bytes = ((row * (size_t)n_bins * 2 * rs + 255) / 256) * 256;
      if (P.ops[s].def_off >= 0) {
        // table indexed by absolute bin: shift the base by -bin_begin rows (never dereferenced below bin_begin)
        char* tbase = tabp - (size_t)bin_begin * row * 2 * rs;
        Pdef.ops[s].gtab = tbase;
        DeferArgs D;
        memset(&D, 0, sizeof(D));
        D.n_bins = n_bins;
        D.bin_begin = bin_begin;
        D.opi = s;
        const int pairs = (int)row;
        const int chunks = (int)((n_bins + DEF_BLOCK * DEF_TILES - 1) / (DEF_BLOCK * DEF_TILES));
        if (plan->dtype == FSWEEP_C64)
          fsweep_sos_table_kernel<float><<<dim3((unsigned)pairs, (unsigned)chunks), DEF_BLOCK, 0, st>>>(Pdef, D);
        else
          fsweep_sos_table_kernel<double><<<dim3((unsigned)pairs, (unsigned)chunks), DEF_BLOCK, 0, st>>>(Pdef, D);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return fail(FSWEEP_E_CUDA, "response table launch: %s", cudaGetErrorString(e2));
        ++launches;
        P.ops[s].kind = par ? FSWEEP_OP_PTABLE : FSWEEP_OP_TABLE;
        P.ops[s].coef = tbase;
      }
      tabp += bytes;
    }
  }
  const int dtype = plan->dtype;
  StreamInfo SI;
  size_t ssmem = 0;
  bool stream_tma = false;
  if (plan->cta) {
    cfg.grid = cta_grid(plan, true, n_bins, &e);
    if (e == cudaSuccess) e = launch_cta(plan->dtype, true, plan->cta_tc, cfg.grid, st, P, plan->loop, A, plan->G);
  } else if (!crit && ((stream_tma = plan->stream && stream_tma_ok(P, plan->sinfo, bin_begin, true) &&
                                     stream_setup(plan, batch * cols, true, true, &SI, &ssmem)) ||
                       stream_setup(plan, batch * cols, true, false, &SI, &ssmem))) {
    int bps = plan->stream_bps_smem[1] == ssmem ? plan->stream_bps[1] : 0;
    if (bps == 0) {
      e = occupancy_stream(true, stream_tma, SI.threads, ssmem, &bps);
      if (e == cudaSuccess) {
        plan->stream_bps[1] = bps;
        plan->stream_bps_smem[1] = ssmem;
      }
    }
    if (e == cudaSuccess) {
      const int64_t tiles = (n_bins + SI.tb - 1) / SI.tb;
      cfg.grid = (int)std::min<int64_t>(std::min<int64_t>(tiles, (int64_t)std::max(1, bps) * std::max(1, plan->num_sms ? plan->num_sms : 148)),
                                        grid_cap(n_bins, plan->G));
      e = launch_stream(true, stream_tma, cfg.grid, ssmem, st, P, SI, A, plan->G);
    }
  } else if (use_tpc(plan, n_bins) && use_tpr(plan)) {
    cfg.grid = tpr_grid(plan, n_bins, &e);
    if (e == cudaSuccess) e = launch_tpr(plan->tpc_np, true, cfg.grid, st, P, plan->loop, A, plan->G);
  } else if (use_tpc(plan, n_bins)) {
    cfg.grid = tpc_grid(n_bins);
    e = launch_tpc(plan->tpc_np, true, cfg.grid, st, P, plan->loop, A, plan->G);
  } else if (use_tpb(plan, n_bins, true)) {
    const LoopInfo L = plan->loop;
    cfg.grid = (int)std::min<int64_t>((n_bins + TPB_BLOCK - 1) / TPB_BLOCK, grid_cap(n_bins, plan->G));
    const size_t smem = (size_t)std::max(1, P.acc_total) * TPB_BLOCK * sizeof(float);
    e = launch_tpb_bwd(plan->tpb_np, cfg.grid, smem, st, P, L, A, plan->G);
  } else if (loop) {
    const LoopInfo L = plan->loop;
    e = by_group(plan->G, [&](auto g) { return launch_loop_bwd<decltype(g)::value>(dtype, cfg, P, L, A); });
  } else {
    e = by_group(plan->G, [&](auto g) { return launch_bwd<decltype(g)::value>(dtype, cc, cfg, P, A); });
  }
  if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "backward launch: %s", cudaGetErrorString(e));
  ++launches;

  if (any_deferred) {
    // absolute bins k with 4k <= nfft use Taylor block 0
    const int64_t k_last_plus = P.nfft / 4;
    const int64_t n_plus = std::max<int64_t>(0, std::min<int64_t>(n_bins, k_last_plus - bin_begin + 1));
    for (int s = 0; s < Pdef.n_ops; ++s) {
      if (Pdef.ops[s].def_off < 0) continue;
      DeferArgs D;
      memset(&D, 0, sizeof(D));
      D.defer = A.defer;
      D.gacc = gacc;
      D.n_bins = n_bins;
      D.bin_begin = bin_begin;
      D.n_plus = n_plus;
      D.ncols_total = (int)(batch * cols);
      D.opi = s;
      const int pairs = Pdef.ops[s].kind == FSWEEP_OP_PSOS ? Pdef.ops[s].n_out : Pdef.ops[s].n_out * Pdef.ops[s].n_in;
      const int ch = DEF_BLOCK * DEF_TILES;
      D.chunks_plus = (int)((n_plus + ch - 1) / ch);
      const int chunks = D.chunks_plus + (int)((n_bins - n_plus + ch - 1) / ch);
      if (dtype == FSWEEP_C64)
        fsweep_sos_defer_kernel<float><<<dim3((unsigned)pairs, (unsigned)chunks), DEF_BLOCK, 0, st>>>(Pdef, D);
      else if (plan->grad32)
        fsweep_sos_defer_kernel<double, float><<<dim3((unsigned)pairs, (unsigned)chunks), DEF_BLOCK, 0, st>>>(Pdef, D);
      else
        fsweep_sos_defer_kernel<double><<<dim3((unsigned)pairs, (unsigned)chunks), DEF_BLOCK, 0, st>>>(Pdef, D);
      if ((e = cudaGetLastError()) != cudaSuccess) return fail(FSWEEP_E_CUDA, "deferred gradient launch: %s", cudaGetErrorString(e));
      ++launches;
    }
  }

  if (any_acc_wanted) {
    FinalizeArgs F;
    memset(&F, 0, sizeof(F));
    if (crit) {
      F.loss_partial = A.loss_partial;
      F.loss = crit->loss;
      F.crit_scale = A.crit_scale;
    }
    F.n_ops = P.n_ops;
    F.G = plan->G;
    F.acc_per_lane = P.acc_per_lane;
    F.n_blocks = cfg.grid;
    F.n_items = (int)items;
    F.acc_total = P.acc_total;
    F.partial = partial;
    F.gacc = gacc;
    int max_total = 1;
    for (int s = 0; s < P.n_ops; ++s) {
      FinalizeOp& o = F.ops[s];
      o.kind = Pdef.ops[s].kind;  // (the sweep kernel may have seen a deferred cascade as a TABLE)
      o.n_out = P.ops[s].n_out;
      o.n_in = P.ops[s].n_in;
      o.K = P.ops[s].K;
      o.acc_mode = P.ops[s].acc_mode;
      o.row_off = P.ops[s].row_off;
      o.row_len = P.ops[s].row_len;
      o.acc_off = P.ops[s].acc_off;
      o.grad = grad_coeffs[s];
      o.grad_is = plan->items && plan->leaf[s].per_item ? (long long)fsweep_plan_coeff_numel(plan, s, P.nfft / 2 + 1) : 0;
      if (o.grad && (o.acc_mode == ACC_SMEM || o.acc_mode == ACC_GLOBAL))
        max_total = std::max(max_total, o.n_out * o.row_len);
    }
    // 4 warps per block, 1 warp per element; row n_ops of the grid sums the fused criterion's loss
    dim3 grid((unsigned)std::min(1024, (max_total + 3) / 4), (unsigned)P.n_ops + (crit ? 1u : 0u), (unsigned)items);
    const char* v2 = getenv("FSWEEP_FINALIZE_V2");  // coalesced kernel (default); "0" selects the round-1 kernel
    if (!(v2 && v2[0] == '0')) {
      dim3 grid2((unsigned)std::min(256, (max_total + 31) / 32), grid.y, grid.z);
      if (dtype == FSWEEP_C64)
        launch_pdl(fsweep_finalize_v2_kernel<float>, grid2, dim3(32 * FIN2_WARPS), 0, st, F);
      else
        launch_pdl(fsweep_finalize_v2_kernel<double>, grid2, dim3(32 * FIN2_WARPS), 0, st, F);
    } else if (dtype == FSWEEP_C64)
      fsweep_finalize_kernel<float><<<grid, 128, 0, st>>>(F);
    else
      fsweep_finalize_kernel<double><<<grid, 128, 0, st>>>(F);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FSWEEP_E_CUDA, "finalize launch: %s", cudaGetErrorString(e));
    ++launches;
  } else if (crit) {
    launch_loss_finalize(dtype, cfg.grid, (int)items, A, crit->loss, st);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(FSWEEP_E_CUDA, "loss finalize launch: %s", cudaGetErrorString(e));
    ++launches;
  }
  g_launches = launches;
  return FSWEEP_OK;
}
}  // namespace
