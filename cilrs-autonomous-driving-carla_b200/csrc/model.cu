// Whole-network plan for the CILRS hot path: ResNet-34 trunk (tcgen05 implicit-GEMM convs + BN/ReLU/pool kernels),
// heads, losses, backward pass. One C-ABI call per forward / backward; all memory is caller-owned:
//   params / grads : one flat fp32 arena each, tensors in the reference's named_parameters() order
//                    (model/autonomous_drive.py:361-399; layout exported by cilrs_model_param_layout)
//   buffers        : flat fp32 arena of BN running_mean / running_var + int64 num_batches_tracked[36]
//   workspace      : activations, packed bf16 weights, BN vectors, gradient ping-pong buffers
#include "conv_params.h"
#include "conv_host.h"
#include "elementwise.cuh"
#include "stem_pool.cuh"
#include "heads_run.cuh"
#include "adam.cuh"
#include <new>
#include <string.h>
#include <stdlib.h>
#include <vector>

namespace cilrs {

static inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct TensorSlot {
  long long off;   // offset in floats inside the arena
  long long size;  // elements
};

struct BnRef {
  int C;
  int gamma, beta;  // param slot indices
  long long rm_off, rv_off;
  int nbt_idx;
  float* vec;   // [4][C] scale, shift, mean, rstd   (workspace)
  float* bred;  // [2][C] bsum, bdot                  (workspace)
  double* acc;  // [2][C] fp64 sums of the fused forward statistics (zeroed at the start of every training forward)
  double* bacc; // [3][C] fp64 sums of the fused backward reductions (zeroed at the start of every backward)
};

struct ConvRef {
  cilrs_conv_desc d;
  bool stem;
  bool flat;  // 3x3 stride 1: runs on the padded-flat kernels (conv_flat.cuh / wgrad_flat.cuh)
  int w;      // param slot
  __nv_bfloat16 *wf, *wd;
  BnRef bn;
  __nv_bfloat16* y;  // raw conv output (pre-BN), padded-flat (the stem's is dense)
  int oh, ow;
  PadGeom gin, gout;  // padded-flat geometry of the input / output tensors
};

// indices into the plan vectors of the model (-1 = not used by this block / mode)
struct BlockPlan {
  int f_a_old = -1, f_a_flat = -1, f_ds_old = -1, f_b_flat = -1;
  int w_b = -1, d_b = -1, w_a_flat = -1, w_a_old = -1, w_ds_old = -1, d_a_flat = -1, d_a_old = -1;  // d_a_old: first of 4 parity plans
};

struct Block {
  ConvRef a, b, ds;
  bool has_ds;
  const __nv_bfloat16* in;
  __nv_bfloat16* act_a;
  __nv_bfloat16* out;
  uint8_t* bits_a;    // ReLU bits of act_a / out (one bit per element), written by bn_apply, read by the flat dgrad epilogues
  uint8_t* bits_out;
  int in_h, in_w, in_c;
  BlockPlan pl;
};

struct Bump {
  char* base;
  long long off;
  void* take(long long bytes) {
    off = align_up(off, 1024);
    void* p = base ? (void*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

// optional per-kernel-class timing (CUDA events around every launch of the plan; used by bench.py's roofline report,
// never inside a timed region)
enum ProfClass { PC_FPROP = 0, PC_DGRAD, PC_WGRAD, PC_BN_FWD, PC_BN_BWD, PC_HEADS, PC_OTHER, PC_COUNT };
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  ~Prof() { for (auto e : pool) cudaEventDestroy(e); }
};
#define PROF(m, cls, s, ...)                        \
  do {                                              \
    if ((m).prof.on) {                              \
      cudaEvent_t _a = (m).prof.get();              \
      cudaEventRecord(_a, s);                       \
      __VA_ARGS__;                                  \
      cudaEvent_t _b = (m).prof.get();              \
      cudaEventRecord(_b, s);                       \
      (m).prof.recs.push_back(Prof::Rec{cls, _a, _b}); \
    } else {                                        \
      __VA_ARGS__;                                  \
    }                                               \
  } while (0)

struct Model {
  Prof prof;
  int maxB;
  std::vector<TensorSlot> slots;  // parameter tensors
  long long param_floats;
  long long buffer_floats;
  int num_bn;
  ConvRef stem;
  std::vector<Block> blocks;
  int head_slot0;  // first slot of the heads
  // bound arenas
  float* params = nullptr;
  float* grads = nullptr;
  float* buffers = nullptr;
  long long* nbt = nullptr;
  // workspace
  __nv_bfloat16* x_s2d;
  __nv_bfloat16* pool_out;
  uint8_t* pool_arg;
  __nv_bfloat16* pool_ysel;   // raw conv1 output at the arg-max of every pool window (padded-flat like pool_out): the stem's BN-backward sums
  float* stats;          // shared stats-partials scratch
  float* stats_ds;       // the same for the downsample convolutions (they run beside conv_a on the gradient stream)
  double* stat_acc;
  double* acc_fwd;       // per-BatchNorm accumulators of the deferred finalize (conv_params.h: CF_DEFER): forward, backward
  double* acc_bwd;
  long long acc_fwd_bytes, acc_bwd_bytes;
  unsigned int *bar_fwd = nullptr, *bar_bwd = nullptr;  // grid-barrier words of the fused launches (CF_FUSE), zeroed with the accumulators
  int bar_fwd_next = 0, bar_bwd_next = 0;
  bool bw_fused_prev = false;  // dy of the current block's conv_b (and downsample) BatchNorms was already written by the fused dgrad of the block above
  bool bw_deferred = false;  // the reductions of the current block-output gradient sit in bn_b.bacc (fused dgrad) rather than in bred
  unsigned int* counters;
  float* unit_vec;       // [2][64]: ones, zeros (max-pool on already-activated stem output)
  float* feat;           // [B,512]
  float* dfeat;
  float* dfeat2;         // feature gradient through the speed predictor (dfeat: through the command branch)
  float* head_comb;      // [B,640] = [feat | sfeat]: input of the first branch layer, staged for its weight gradient
  HeadsSaved hs;
  float* loss_out;       // [8]
  float* dcontrols;      // [B,3]
  float* dspeed;         // [B]
  int* err_flag;
  const long long* drop_counter = nullptr;  // optional device counter mixed into the dropout seed (CUDA-graph replays)
  __nv_bfloat16 *g0, *g1, *ga, *dy_stem;
  // dy ring: the weight-gradient kernels run on a side stream concurrently with the dgrad / BatchNorm chain, so the buffer
  // a wgrad reads must not be overwritten before it finished: block bi uses dyb[bi & 1] (conv_b), dya[bi & 1] (conv_a),
  // dyd[k & 1] (k-th downsample)
  __nv_bfloat16 *dyb[2], *dya[2], *dyd[2];
  cudaStream_t side = nullptr;
  cudaEvent_t ev_part = nullptr;   // end of an asynchronous backward part on the caller's stream
  cudaEvent_t ev_heads = nullptr;  // head deltas ready: the head weight gradients run on the side stream
  cudaEvent_t ev_ready[6] = {}, ev_done[6] = {}, ev_join = nullptr;  // slot order: dyb0, dyb1, dya0, dya1, dyd0, dyd1
  cudaEvent_t ev_pack_fork = nullptr, ev_pack = nullptr;             // cilrs_model_refresh_async: repack beside the stem
  bool pack_pending = false;
  bool pending[6] = {false, false, false, false, false, false};      // a side-stream wgrad still reads the slot
  double* sumsq_partial;
  PackJob* pack_jobs;    // device table for the one-launch weight repack
  int pack_njobs = 0, pack_blocks = 0;
  int pack_block_first[17] = {};  // first pack-kernel block of BasicBlock bi's jobs ([16] = total): repack of one backward part
  // plans (tensor maps) for the batch size they were built for
  int planB = 0;
  int planMode = -1;
  std::vector<ConvGemmParams> old_plans;     // generic kernel: stem, stride-2 and 1x1 convs (fprop and dgrad)
  std::vector<WgradParams> wold_plans;
  std::vector<FlatConvParams> flat_plans;    // padded-flat kernel: 3x3 stride-1 fprop and dgrad
  std::vector<WgradFlatParams> wflat_plans;
  std::vector<long long> wflat_grad_off;     // gradient-arena offset each flat wgrad plan reduces into
  WgradReduceJobs red_jobs[4];               // split-K reductions of the flat wgrads, one launch per backward part (layer4..layer1)
  float* wscratch = nullptr;                 // WF_SCRATCH_BYTES per flat 3x3 convolution
  int stem_fwd = -1, stem_wgrad = -1;
  // resumable backward (so the host can start the allreduce of finished gradient buckets between parts)
  __nv_bfloat16 *bw_gcur = nullptr, *bw_gnext = nullptr;
};

static int add_slot(Model& m, long long size) {
  TensorSlot s;
  s.off = align_up(m.param_floats, 16);
  s.size = size;
  m.param_floats = s.off + size;
  m.slots.push_back(s);
  return (int)m.slots.size() - 1;
}

static BnRef make_bn(Model& m, int C) {
  BnRef b{};
  b.C = C;
  b.gamma = add_slot(m, C);
  b.beta = add_slot(m, C);
  b.rm_off = m.buffer_floats;
  b.rv_off = m.buffer_floats + C;
  m.buffer_floats += 2 * C;
  b.nbt_idx = m.num_bn++;
  return b;
}

static ConvRef make_conv(Model& m, int B, int h, int w, int cin, int cout, int k, int stride) {
  ConvRef c{};
  c.stem = false;
  c.flat = (k == 3 && stride == 1);
  c.d = cilrs_conv_desc{B, h, w, cin, cout, k, k, stride, k == 3 ? 1 : 0};
  c.w = add_slot(m, (long long)cout * cin * k * k);
  c.bn = make_bn(m, cout);
  c.oh = conv_out_dim(h, k, stride, c.d.pad);
  c.ow = conv_out_dim(w, k, stride, c.d.pad);
  c.gin = PadGeom{h, w, h + 1, w + 1};
  c.gout = PadGeom{c.oh, c.ow, c.oh + 1, c.ow + 1};
  return c;
}

// network topology; also the flat parameter layout (named_parameters order of the reference module)
static void build_topology(Model& m, int B) {
  m.param_floats = 0; m.buffer_floats = 0; m.num_bn = 0;
  m.slots.clear(); m.blocks.clear();
  m.stem = ConvRef{};
  m.stem.stem = true;
  m.stem.d = cilrs_conv_desc{B, 88, 200, 3, 64, 7, 7, 2, 3};
  m.stem.w = add_slot(m, 64 * 3 * 7 * 7);
  m.stem.bn = make_bn(m, 64);
  m.stem.oh = 44; m.stem.ow = 100;
  const int chans[4] = {64, 128, 256, 512}, nblk[4] = {3, 4, 6, 3};
  int h = 22, w = 50, cin = 64;
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < nblk[s]; ++b) {
      Block blk{};
      const int stride = (b == 0 && s > 0) ? 2 : 1;
      const int c = chans[s];
      blk.in_h = h; blk.in_w = w; blk.in_c = cin;
      blk.a = make_conv(m, B, h, w, cin, c, 3, stride);
      blk.b = make_conv(m, B, blk.a.oh, blk.a.ow, c, c, 3, 1);
      blk.has_ds = (stride != 1 || cin != c);
      if (blk.has_ds) blk.ds = make_conv(m, B, h, w, cin, c, 1, stride);
      h = blk.a.oh; w = blk.a.ow; cin = c;
      m.blocks.push_back(blk);
    }
  }
  m.head_slot0 = (int)m.slots.size();
  add_slot(m, 128); add_slot(m, 128);          // speed_encoder.0 weight [128,1], bias
  add_slot(m, 128 * 128); add_slot(m, 128);    // speed_encoder.3
  for (int k = 0; k < 4; ++k) {
    add_slot(m, 256 * 640); add_slot(m, 256);  // control_branches.k.0
    add_slot(m, 256 * 256); add_slot(m, 256);  // .3
    add_slot(m, 3 * 256); add_slot(m, 3);      // .6
  }
  add_slot(m, 256 * 512); add_slot(m, 256);    // speed_predictor.0
  add_slot(m, 256 * 256); add_slot(m, 256);    // .3
  add_slot(m, 256); add_slot(m, 1);            // .5
  m.param_floats = align_up(m.param_floats, 16);
}

static long long act_elems(int B, int h, int w, int c) { return (long long)B * h * w * c; }
static long long pad_elems(int B, const PadGeom& g, int c) { return (long long)B * g.Hp * g.Wp * c; }
static const PadGeom kGeom0{22, 50, 23, 51};   // layer1 / max-pool output
static const PadGeom kDense{1, 1, 1, 1};       // "every pixel is real" for the elementwise kernels

static void carve_conv(Bump& bp, ConvRef& c, int B) {
  const long long wbytes = c.stem ? 4 * 64 * 64 * 2 : (long long)c.d.kh * c.d.kw * c.d.in_c * c.d.out_c * 2;
  c.wf = (__nv_bfloat16*)bp.take(wbytes);
  c.wd = c.stem ? nullptr : (__nv_bfloat16*)bp.take(wbytes);
  c.y = (__nv_bfloat16*)bp.take((c.stem ? act_elems(B, c.oh, c.ow, c.d.out_c) : pad_elems(B, c.gout, c.d.out_c)) * 2);
  c.bn.vec = (float*)bp.take(4LL * c.bn.C * 4);
  c.bn.bred = (float*)bp.take(2LL * c.bn.C * 4);
}

static long long carve(Model& m, char* base) {
  Bump bp{base, 0};
  const int B = m.maxB;
  m.x_s2d = (__nv_bfloat16*)bp.take((long long)B * 47 * 103 * 16 * 2);
  carve_conv(bp, m.stem, B);
  m.pool_out = (__nv_bfloat16*)bp.take(pad_elems(B, kGeom0, 64) * 2);
  m.pool_arg = (uint8_t*)bp.take(act_elems(B, 22, 50, 64));
  m.pool_ysel = (__nv_bfloat16*)bp.take(pad_elems(B, kGeom0, 64) * 2);
  const __nv_bfloat16* prev = m.pool_out;
  for (auto& blk : m.blocks) {
    blk.in = prev;
    carve_conv(bp, blk.a, B);
    carve_conv(bp, blk.b, B);
    if (blk.has_ds) carve_conv(bp, blk.ds, B);
    blk.act_a = (__nv_bfloat16*)bp.take(pad_elems(B, blk.a.gout, blk.a.d.out_c) * 2);
    blk.out = (__nv_bfloat16*)bp.take(pad_elems(B, blk.b.gout, blk.b.d.out_c) * 2);
    blk.bits_a = (uint8_t*)bp.take(pad_elems(B, blk.a.gout, blk.a.d.out_c) / 8);
    blk.bits_out = (uint8_t*)bp.take(pad_elems(B, blk.b.gout, blk.b.d.out_c) / 8);
    prev = blk.out;
  }
  // stats partial scratch: the stem has the most tiles (<= ceil(B*4400/100) ~ 44*B + slack), 2 x 64 floats each;
  // deeper layers have fewer tiles x more channels; bound by B*44*100/64 tiles * 2 * 64
  m.stats_ds = (float*)bp.take(256LL * 2 * 512 * 4);
  m.stats = (float*)bp.take(256LL * 2 * 512 * 4);  // one (sum, sumsq)[C<=512] partial per persistent conv CTA (<= SM count)
  m.stat_acc = (double*)bp.take(3 * 512 * 8);  // per-channel fp64 statistics accumulators (kept zero between launches)
  {
    long long ch = m.stem.bn.C;
    for (auto& blk : m.blocks) ch += blk.a.bn.C + blk.b.bn.C + (blk.has_ds ? blk.ds.bn.C : 0);
    m.acc_fwd_bytes = ch * 2 * 8 + 64 * 4; m.acc_bwd_bytes = ch * 3 * 8 + 64 * 4;   // + 64 grid-barrier words each
    m.acc_fwd = (double*)bp.take(m.acc_fwd_bytes);
    m.acc_bwd = (double*)bp.take(m.acc_bwd_bytes);
    m.bar_fwd = m.acc_fwd ? (unsigned int*)(m.acc_fwd + ch * 2) : nullptr;
    m.bar_bwd = m.acc_bwd ? (unsigned int*)(m.acc_bwd + ch * 3) : nullptr;
    long long off = 0;
    auto give = [&](BnRef& bn) {
      bn.acc = m.acc_fwd ? m.acc_fwd + 2 * off : nullptr;
      bn.bacc = m.acc_bwd ? m.acc_bwd + 3 * off : nullptr;
      off += bn.C;
    };
    give(m.stem.bn);
    for (auto& blk : m.blocks) { give(blk.a.bn); give(blk.b.bn); if (blk.has_ds) give(blk.ds.bn); }
  }
  m.counters = (unsigned int*)bp.take(64);
  m.unit_vec = (float*)bp.take(2 * 64 * 4);
  m.feat = (float*)bp.take((long long)B * 512 * 4);
  m.dfeat = (float*)bp.take((long long)B * 512 * 4);
  m.dfeat2 = (float*)bp.take((long long)B * 512 * 4);
  m.head_comb = (float*)bp.take((long long)B * 640 * 4);
  m.hs.s1 = (float*)bp.take((long long)B * 128 * 4);
  m.hs.sfeat = (float*)bp.take((long long)B * 128 * 4);
  m.hs.b1 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.b2 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.p1 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.p2 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.d_se0 = (float*)bp.take((long long)B * 128 * 4);
  m.hs.d_se3 = (float*)bp.take((long long)B * 128 * 4);
  m.hs.d_br0 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.d_br3 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.d_br6 = (float*)bp.take((long long)B * 4 * 4);
  m.hs.d_sp0 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.d_sp3 = (float*)bp.take((long long)B * 256 * 4);
  m.hs.d_sp5 = (float*)bp.take((long long)B * 4);
  m.loss_out = (float*)bp.take(64);
  m.dcontrols = (float*)bp.take((long long)B * 3 * 4);
  m.dspeed = (float*)bp.take((long long)B * 4);
  m.err_flag = (int*)bp.take(64);
  const long long gmax = pad_elems(B, kGeom0, 64) * 2;  // the largest padded-flat activation (layer1)
  m.g0 = (__nv_bfloat16*)bp.take(gmax);
  m.g1 = (__nv_bfloat16*)bp.take(gmax);
  m.ga = (__nv_bfloat16*)bp.take(gmax);
  for (int i = 0; i < 2; ++i) {
    m.dyb[i] = (__nv_bfloat16*)bp.take(gmax);
    m.dya[i] = (__nv_bfloat16*)bp.take(gmax);
    m.dyd[i] = (__nv_bfloat16*)bp.take(gmax);
  }
  m.dy_stem = (__nv_bfloat16*)bp.take(act_elems(B, 44, 100, 64) * 2);
  m.sumsq_partial = (double*)bp.take(1024 * 8);
  m.wscratch = (float*)bp.take(29LL * WF_SCRATCH_BYTES);
  m.pack_jobs = (PackJob*)bp.take(64 * sizeof(PackJob));
  return align_up(bp.off, 1024);
}

static void set_batch(ConvRef& c, int B) { c.d.batch = B; }

// ------------------------------------------------------------------------------------------------
// plan construction: every tensor map for batch B (forward variants, dgrad, wgrad)
// ------------------------------------------------------------------------------------------------
enum FwdMode { MODE_TRAIN = 0, MODE_FROZEN = 1, MODE_INFER = 2 };

#define CK(call)            \
  do {                      \
    int _st = (call);       \
    if (_st) return _st;    \
  } while (0)
#define CKL() CK(cuda_status(cudaGetLastError()))

// launch-geometry knobs of the stem's elementwise kernels and the stand-alone reduce (CTAs per SM; measurement aids)
static int env_cap(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v >= 1 ? v : dflt;
}
static int capped_grid(long long blocks, int per_sm) {
  const long long cap = 148LL * per_sm;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

static int add_old(Model& m, const ConvGemmParams& p) { m.old_plans.push_back(p); return (int)m.old_plans.size() - 1; }
static int add_flat(Model& m, const FlatConvParams& p) { m.flat_plans.push_back(p); return (int)m.flat_plans.size() - 1; }

static int build_plans(Model& m, int B, int mode) {
  if (B == m.planB && mode == m.planMode) return OK;
  m.old_plans.clear(); m.wold_plans.clear(); m.flat_plans.clear(); m.wflat_plans.clear(); m.wflat_grad_off.clear();
  for (int i = 0; i < 4; ++i) { m.red_jobs[i].n = 0; m.red_jobs[i].total_blocks = 0; }
  set_batch(m.stem, B);
  const bool infer = mode == MODE_INFER;
  const bool training = mode == MODE_TRAIN;
  {
    ConvGemmParams p;
    BnRef& bn = m.stem.bn;
    if (infer) CK(build_stem_fprop(&p, B, m.x_s2d, m.stem.wf, m.stem.y, bn.vec, bn.vec + bn.C, nullptr, CG_SCALE_BIAS | CG_RELU));
    else CK(build_stem_fprop(&p, B, m.x_s2d, m.stem.wf, m.stem.y, nullptr, nullptr, m.stats, CG_STATS));
    m.stem_fwd = add_old(m, p);
  }
  for (auto& blk : m.blocks) {
    set_batch(blk.a, B); set_batch(blk.b, B);
    if (blk.has_ds) set_batch(blk.ds, B);
    blk.pl = BlockPlan{};
    ConvGemmParams p;
    FlatConvParams f;
    // ---- forward ----
    // conv_a: inference fuses folded BN + ReLU and writes act_a directly; training writes the raw output (+ BN statistics)
    if (blk.a.flat) {
      const int fl = infer ? (CF_SCALE_BIAS | CF_RELU) : (training ? CF_STATS : 0);
      CK(build_flat_conv(&f, B, blk.a.gin, blk.a.d.in_c, blk.a.d.out_c, 0, blk.in, blk.a.wf, infer ? blk.act_a : blk.a.y, fl));
      if (infer) { f.scale = blk.a.bn.vec; f.bias = blk.a.bn.vec + blk.a.bn.C; }
      if (training && flat_conv_fuse_ok(&f)) {   // BatchNorm + ReLU applied by the conv's own second pass (CF_FUSE)
        f.flags |= CF_FUSE;
        f.bits_out = blk.bits_a;
        CK(flat_conv_bind_fuse(&f, blk.act_a, nullptr));
      }
      blk.pl.f_a_flat = add_flat(m, f);
    } else {
      if (infer) CK(build_fprop(&p, &blk.a.d, blk.in, blk.a.wf, blk.act_a, blk.a.bn.vec, blk.a.bn.vec + blk.a.bn.C, nullptr, nullptr,
                                CG_SCALE_BIAS | CG_RELU, &blk.a.gin, &blk.a.gout));
      else CK(build_fprop(&p, &blk.a.d, blk.in, blk.a.wf, blk.a.y, nullptr, nullptr, nullptr, m.stats, CG_STATS, &blk.a.gin, &blk.a.gout));
      blk.pl.f_a_old = add_old(m, p);
    }
    if (blk.has_ds) {
      if (infer) CK(build_fprop(&p, &blk.ds.d, blk.in, blk.ds.wf, blk.ds.y, blk.ds.bn.vec, blk.ds.bn.vec + blk.ds.bn.C, nullptr, nullptr,
                                CG_SCALE_BIAS, &blk.ds.gin, &blk.ds.gout));
      else CK(build_fprop(&p, &blk.ds.d, blk.in, blk.ds.wf, blk.ds.y, nullptr, nullptr, nullptr, m.stats_ds, CG_STATS, &blk.ds.gin, &blk.ds.gout));
      blk.pl.f_ds_old = add_old(m, p);
    }
    {
      const int fl = infer ? (CF_SCALE_BIAS | CF_RESIDUAL | CF_RELU) : (training ? CF_STATS : 0);
      CK(build_flat_conv(&f, B, blk.b.gin, blk.b.d.in_c, blk.b.d.out_c, 0, blk.act_a, blk.b.wf, infer ? blk.out : blk.b.y, fl));
      if (infer) {
        f.scale = blk.b.bn.vec; f.bias = blk.b.bn.vec + blk.b.bn.C;
        f.residual = blk.has_ds ? blk.ds.y : blk.in;
        CK(flat_conv_bind_operands(&f));
      }
      if (training && flat_conv_fuse_ok(&f)) {   // out = relu(bn_b(y_b) + identity | bn_ds(y_ds)) by the conv's second pass
        f.flags |= CF_FUSE | CF_FUSE_RES;
        f.residual = blk.has_ds ? blk.ds.y : blk.in;
        f.fuse_rvec = blk.has_ds ? blk.ds.bn.vec : nullptr;
        f.bits_out = blk.bits_out;
        CK(flat_conv_bind_operands(&f));
        CK(flat_conv_bind_fuse(&f, blk.out, nullptr));
      }
      blk.pl.f_b_flat = add_flat(m, f);
    }
  }
  if (!infer) {
    // backward plans. Gradient buffers ping-pong g0/g1 (gcur = masked gradient dz of the current block's output).
    __nv_bfloat16* gcur = m.g0;
    __nv_bfloat16* gnext = m.g1;
    for (int bi = (int)m.blocks.size() - 1; bi >= 0; --bi) {
      Block& blk = m.blocks[bi];
      ConvGemmParams p;
      FlatConvParams f;
      WgradParams wp;
      WgradFlatParams wf;
      // conv_b: dW_b = wgrad(dy_b = d1, act_a);  ga = relu'(act_a) * dgrad_b(d1), + the BN_a backward reductions
      const int part = bi >= 13 ? 0 : (bi >= 7 ? 1 : (bi >= 3 ? 2 : 3));
      __nv_bfloat16* dyb = m.dyb[bi & 1];
      __nv_bfloat16* dya = m.dya[bi & 1];
      __nv_bfloat16* dyd = m.dyd[part & 1];  // one downsample per layer group
      auto add_wflat = [&](const ConvRef& c, const __nv_bfloat16* dy, const __nv_bfloat16* x, int* idx) -> int {
        float* scratch = (float*)((char*)m.wscratch + (long long)m.wflat_plans.size() * WF_SCRATCH_BYTES);
        CK(build_wgrad_flat(&wf, B, c.gin, c.d.in_c, c.d.out_c, dy, x, scratch));
        m.wflat_plans.push_back(wf);
        *idx = (int)m.wflat_plans.size() - 1;
        return add_wgrad_reduce_job(&m.red_jobs[part], &wf, m.slots[c.w].off);
      };
      CK(add_wflat(blk.b, dyb, blk.act_a, &blk.pl.w_b));
      CK(build_flat_conv(&f, B, blk.b.gin, blk.b.d.out_c, blk.b.d.in_c, 1, dyb, blk.b.wd, m.ga, CF_MASK | CF_BNBWD));
      f.mask = blk.act_a; f.mask_bits = blk.bits_a; f.y1 = blk.a.y; f.stat1 = blk.a.bn.vec; f.bred1 = blk.a.bn.bred;
      CK(flat_conv_bind_operands(&f));
      if (training && flat_conv_fuse_ok(&f)) {   // dy_a straight from the dgrad's accumulators: dz_a is never stored
        f.flags |= CF_FUSE | CF_NO_STORE;
        CK(flat_conv_bind_fuse(&f, dya, nullptr));
      }
      blk.pl.d_b = add_flat(m, f);
      // conv_a: dW_a = wgrad(dy_a = d1, in); downsample: dW_ds = wgrad(dy_ds = d2, in)
      if (blk.a.flat) {
        CK(add_wflat(blk.a, dya, blk.in, &blk.pl.w_a_flat));
      } else {
        CK(build_wgrad(&wp, &blk.a.d, dya, blk.in, nullptr, &blk.a.gin, &blk.a.gout));
        m.wold_plans.push_back(wp); blk.pl.w_a_old = (int)m.wold_plans.size() - 1;
      }
      if (blk.has_ds) {
        CK(build_wgrad(&wp, &blk.ds.d, dyd, blk.in, nullptr, &blk.ds.gin, &blk.ds.gout));
        m.wold_plans.push_back(wp); blk.pl.w_ds_old = (int)m.wold_plans.size() - 1;
      }
      // gradient of the block input: dgrad_a(d1) + identity path (dz of this block) | + dgrad_ds(d2)
      if (blk.a.flat) {
        int fl = CF_RESIDUAL;
        if (bi > 0) fl |= CF_MASK | CF_BNBWD | (m.blocks[bi - 1].has_ds ? CF_BNBWD2 : 0);
        CK(build_flat_conv(&f, B, blk.a.gin, blk.a.d.out_c, blk.a.d.in_c, 1, dya, blk.a.wd, gnext, fl));
        f.residual = gcur;
        if (bi > 0) {
          Block& pb = m.blocks[bi - 1];
          f.mask = blk.in;  // = output of the previous block
          f.mask_bits = pb.bits_out;
          f.y1 = pb.b.y; f.stat1 = pb.b.bn.vec; f.bred1 = pb.b.bn.bred;
          if (pb.has_ds) { f.y2 = pb.ds.y; f.stat2 = pb.ds.bn.vec; f.bred2 = pb.ds.bn.bred; }
        }
        CK(flat_conv_bind_operands(&f));
        if (training && bi > 0 && flat_conv_fuse_ok(&f)) {   // dy of the block below's conv_b (and downsample) BatchNorms by the second pass
          Block& pb = m.blocks[bi - 1];
          const int pb_part = (bi - 1) >= 13 ? 0 : ((bi - 1) >= 7 ? 1 : ((bi - 1) >= 3 ? 2 : 3));
          f.flags |= CF_FUSE;
          CK(flat_conv_bind_fuse(&f, m.dyb[(bi - 1) & 1], pb.has_ds ? m.dyd[pb_part & 1] : nullptr));
        }
        blk.pl.d_a_flat = add_flat(m, f);
      } else {
        for (int ph = 0; ph < 2; ++ph)
          for (int pw = 0; pw < 2; ++pw) {
            const bool fuse = (ph == 0 && pw == 0);
            CK(build_dgrad(&p, &blk.a.d, ph, pw, dya, blk.a.wd, gnext, nullptr, fuse ? dyd : nullptr, fuse ? blk.ds.wd : nullptr,
                           &blk.a.gin, &blk.a.gout));
            const int idx = add_old(m, p);
            if (fuse) blk.pl.d_a_old = idx;
          }
      }
      __nv_bfloat16* t = gcur; gcur = gnext; gnext = t;
    }
    WgradParams wp;
    CK(build_stem_wgrad(&wp, B, m.dy_stem, m.x_s2d, nullptr));
    m.wold_plans.push_back(wp); m.stem_wgrad = (int)m.wold_plans.size() - 1;
  }
  m.planB = B;
  m.planMode = mode;
  return OK;
}

static int run_bn_finalize(Model& m, const BnRef& bn, int tiles, double count, int training, int update, cudaStream_t s,
                           const float* stats = nullptr) {
  BnVectors v{bn.vec, bn.vec + bn.C, bn.vec + 2 * bn.C, bn.vec + 3 * bn.C};
  ++g_cilrs_launches;
  return cuda_status(launch_pdl(bn_finalize_kernel, dim3((bn.C + 31) / 32), dim3(1024), 0, s, stats ? stats : (const float*)m.stats, tiles, bn.C, count,
                                (const float*)(m.params + m.slots[bn.gamma].off), (const float*)(m.params + m.slots[bn.beta].off),
                                m.buffers + bn.rm_off, m.buffers + bn.rv_off, m.nbt ? m.nbt + bn.nbt_idx : (long long*)nullptr, 0.1f, 1e-5f,
                                training, update, v));
}

// out = relu?( bn(x) [+ res] [+ bn2(x2)] ) on padded-flat tensors of geometry g
// deferred != 0: the statistics of `bn` are still the raw sums the fused conv epilogue left in bn.acc (training forward)
static int run_bn_apply(Model& m, int B, const PadGeom& g, const __nv_bfloat16* x, const BnRef& bn, const __nv_bfloat16* res,
                        const __nv_bfloat16* x2, const BnRef* bn2, __nv_bfloat16* out, int relu, uint8_t* bits, int deferred,
                        int update_running, cudaStream_t s) {
  const long long nvec = pad_elems(B, g, bn.C) / 8;
  BnDefer d{};
  if (deferred) {
    const double count = (double)B * g.H * g.W;
    d.acc = bn.acc; d.gamma = m.params + m.slots[bn.gamma].off; d.beta = m.params + m.slots[bn.beta].off;
    d.running_mean = m.buffers + bn.rm_off; d.running_var = m.buffers + bn.rv_off; d.nbt = m.nbt ? m.nbt + bn.nbt_idx : nullptr;
    d.vec = bn.vec; d.inv_count = 1.0 / count; d.unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
    d.momentum = 0.1f; d.eps = 1e-5f; d.update_running = update_running;
  }
  ++g_cilrs_launches;
  // CTAs per SM the grid is capped at. Three are resident (70 registers), but measured in the step at batch 128 TWO is the
  // sweet spot: 2.636 ms per step with 3, 2.585 with 2, 2.66 with 1 or 6 (every CTA pays the deferred finalize in its prologue,
  // and the CTAs of the convolution it follows are still draining when it starts). CILRS_EW_FWD_CAP overrides.
  static int cap_f = -1;
  if (cap_f < 0) { const char* e = getenv("CILRS_EW_FWD_CAP"); cap_f = e ? atoi(e) : 2; if (cap_f < 1) cap_f = 2; }
  return cuda_status(launch_pdl(bn_apply_kernel, dim3(ew_grid(nvec, bn.C, 4, cap_f)), dim3(EW_THREADS), 0, s, x, (const float*)bn.vec,
                                (const float*)(bn.vec + bn.C), res, x2, (const float*)(bn2 ? bn2->vec : nullptr),
                                (const float*)(bn2 ? bn2->vec + bn2->C : nullptr), out, nvec, bn.C, relu, g, bits, d));
}

// flat conv with the train-mode BatchNorm statistics + finalize fused (pointers bound at launch time)
static int launch_flat_fwd(Model& m, int idx, const BnRef& bn, double count, int update_running, cudaStream_t s) {
  FlatConvParams f = m.flat_plans[idx];
  if (f.flags & CF_STATS) {
    f.flags |= CF_DEFER;  // the bn_apply that follows finalizes (run_bn_apply, deferred)
    f.partials = bn.acc; f.counter = nullptr;
    f.gamma = m.params + m.slots[bn.gamma].off; f.beta = m.params + m.slots[bn.beta].off;
    f.running_mean = m.buffers + bn.rm_off; f.running_var = m.buffers + bn.rv_off;
    f.nbt = m.nbt ? m.nbt + bn.nbt_idx : nullptr;
    f.vec = bn.vec; f.count = count; f.momentum = 0.1f; f.eps = 1e-5f; f.update_running = update_running;
    if (f.flags & CF_FUSE) {
      if (m.bar_fwd_next >= 64) return ERR_INVALID;
      f.grid_bar = m.bar_fwd + m.bar_fwd_next++;
      f.inv_count = 1.0 / count; f.unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
    }
  }
  return launch_flat_conv(&f, s);
}

static HeadsCtx heads_ctx(const Model& m) {
  HeadsCtx c;
  c.params = m.params; c.grads = m.grads;
  for (int i = 0; i < HD_NUM_SLOTS; ++i) c.off[i] = m.slots[m.head_slot0 + i].off;
  c.feat = m.feat; c.dfeat = m.dfeat; c.dfeat2 = m.dfeat2; c.head_comb = m.head_comb; c.hs = m.hs; c.err_flag = m.err_flag;
  c.drop_counter = m.drop_counter; c.loss_counter = m.counters + 8;
  return c;
}

// re-derive everything that depends on parameter values: packed bf16 conv weights (+ eval-mode BN folding)
// the repack queued by cilrs_model_refresh_async becomes visible to stream s
static int pack_join(Model& m, cudaStream_t s) {
  if (m.pack_pending) {
    CK(cuda_status(cudaStreamWaitEvent(s, m.ev_pack, 0)));
    m.pack_pending = false;
  }
  return OK;
}

static int refresh(Model& m, int what, cudaStream_t s) {
  if (!m.params) return ERR_INVALID;
  CK(pack_join(m, s));
  if (what & 1) {
    CK(cilrs_stem_pack_weight(m.params + m.slots[m.stem.w].off, m.stem.wf, s));
    CK(launch_pack_all(m.params, m.pack_jobs, m.pack_njobs, m.pack_blocks, s));
  }
  if (what & 2) {
    CK(run_bn_finalize(m, m.stem.bn, 0, 1.0, 0, 0, s));
    for (auto& blk : m.blocks) {
      CK(run_bn_finalize(m, blk.a.bn, 0, 1.0, 0, 0, s));
      CK(run_bn_finalize(m, blk.b.bn, 0, 1.0, 0, 0, s));
      if (blk.has_ds) CK(run_bn_finalize(m, blk.ds.bn, 0, 1.0, 0, 0, s));
    }
  }
  return OK;
}

static int heads_forward(Model& m, int B, const float* speed, const long long* command, float* controls, float* pred_speed,
                         int keep_for_backward, float dropout_p, unsigned long long seed, cudaStream_t s, const HeadsLossArgs* loss = nullptr);

static int forward(Model& m, int B, int mode, const float* image, const void* x_s2d_in, const float* speed, const long long* command,
                   float* controls, float* pred_speed, int update_running, int keep_for_backward, float dropout_p,
                   unsigned long long seed, cudaStream_t s, const HeadsLossArgs* loss = nullptr) {
  if (B < 1 || B > m.maxB) return ERR_INVALID;
  if (!m.params || !m.buffers) return ERR_INVALID;
  CK(build_plans(m, B, mode));
  if (image) {
    CK(cilrs_image_to_s2d(image, B, m.x_s2d, s));
  } else if (x_s2d_in) {
    if (x_s2d_in != (const void*)m.x_s2d)
      CK(cuda_status(cudaMemcpyAsync(m.x_s2d, x_s2d_in, (size_t)B * 47 * 103 * 16 * 2, cudaMemcpyDeviceToDevice, s)));
  } else {
    return ERR_INVALID;
  }
  const int training = mode == MODE_TRAIN;
  if (training) CK(cuda_status(cudaMemsetAsync(m.acc_fwd, 0, (size_t)m.acc_fwd_bytes, s)));  // accumulators of the deferred BN finalize (+ grid-barrier words)
  m.bar_fwd_next = 0;
  const long long pool_vec = act_elems(B, 22, 50, 64) / 8;
  PROF(m, PC_FPROP, s, CK(launch_conv_gemm(&m.old_plans[m.stem_fwd], s)));
  if (mode == MODE_INFER) {
    bn_relu_maxpool_sel_kernel<<<ew_grid(pool_vec, 64), EW_THREADS, 0, s>>>(m.stem.y, m.unit_vec, m.unit_vec + 64, m.pool_out, nullptr, nullptr,
                                                                            B, 44, 100, 64, 22, 50, kGeom0.Hp, kGeom0.Wp); ++g_cilrs_launches;
    CKL();
    CK(pack_join(m, s));
    for (auto& blk : m.blocks) {
      if (blk.pl.f_a_flat >= 0) PROF(m, PC_FPROP, s, CK(launch_flat_conv(&m.flat_plans[blk.pl.f_a_flat], s)));
      else PROF(m, PC_FPROP, s, CK(launch_conv_gemm(&m.old_plans[blk.pl.f_a_old], s)));
      if (blk.has_ds) PROF(m, PC_FPROP, s, CK(launch_conv_gemm(&m.old_plans[blk.pl.f_ds_old], s)));
      PROF(m, PC_FPROP, s, CK(launch_flat_conv(&m.flat_plans[blk.pl.f_b_flat], s)));
    }
  } else {
    PROF(m, PC_BN_FWD, s, CK(run_bn_finalize(m, m.stem.bn, conv_gemm_grid(&m.old_plans[m.stem_fwd]), (double)B * 44 * 100, training,
                                             update_running, s)));
    static const int pool_cap = env_cap("CILRS_EW_POOL_CAP", 4);
    bn_relu_maxpool_sel_kernel<<<capped_grid(ew_grid(pool_vec, 64), pool_cap), EW_THREADS, 0, s>>>(m.stem.y, m.stem.bn.vec, m.stem.bn.vec + 64, m.pool_out, m.pool_arg,
                                                                            m.pool_ysel, B, 44, 100, 64, 22, 50, kGeom0.Hp, kGeom0.Wp); ++g_cilrs_launches;
    CKL();
    CK(pack_join(m, s));   // the trunk's operands may have been repacked beside the stem (cilrs_model_refresh_async)
    for (auto& blk : m.blocks) {
      // raw conv output + BN vectors: fused in the flat kernel (training), or generic kernel + finalize kernel
      auto conv_bn = [&](ConvRef& c, int flat_idx, int old_idx, cudaStream_t s, const float* stats = nullptr) -> int {
        const double count = (double)B * c.oh * c.ow;
        if (flat_idx >= 0) {
          PROF(m, PC_FPROP, s, CK(launch_flat_fwd(m, flat_idx, c.bn, count, update_running, s)));
          if (!training) PROF(m, PC_BN_FWD, s, CK(run_bn_finalize(m, c.bn, 0, count, 0, 0, s)));
        } else {
          PROF(m, PC_FPROP, s, CK(launch_conv_gemm(&m.old_plans[old_idx], s)));
          PROF(m, PC_BN_FWD, s, CK(run_bn_finalize(m, c.bn, conv_gemm_grid(&m.old_plans[old_idx]), count, training, update_running, s, stats)));
        }
        return OK;
      };
      // The 1x1 stride-2 downsample convolution (+ its finalize) depends only on the block's input: it runs beside conv_a on the
      // gradient stream (idle during the forward) and is joined before the block's output BatchNorm reads it.
      const bool ds_side = blk.has_ds && m.side != nullptr && !m.prof.on;
      if (ds_side) {
        CK(cuda_status(cudaEventRecord(m.ev_pack_fork, s)));
        CK(cuda_status(cudaStreamWaitEvent(m.side, m.ev_pack_fork, 0)));
        CK(conv_bn(blk.ds, -1, blk.pl.f_ds_old, m.side, m.stats_ds));
        CK(cuda_status(cudaEventRecord(m.ev_pack, m.side)));
      }
      CK(conv_bn(blk.a, blk.pl.f_a_flat, blk.pl.f_a_old, s));
      const int def_a = training && blk.pl.f_a_flat >= 0, def_b = training;
      const bool fused_a = blk.pl.f_a_flat >= 0 && (m.flat_plans[blk.pl.f_a_flat].flags & CF_FUSE);
      const bool fused_b = (m.flat_plans[blk.pl.f_b_flat].flags & CF_FUSE) != 0;
      if (!fused_a)
        PROF(m, PC_BN_FWD, s, CK(run_bn_apply(m, B, blk.a.gout, blk.a.y, blk.a.bn, nullptr, nullptr, nullptr, blk.act_a, 1, blk.bits_a, def_a,
                                              update_running, s)));
      if (blk.has_ds && !ds_side) CK(conv_bn(blk.ds, -1, blk.pl.f_ds_old, s, m.stats_ds));
      if (ds_side && (m.flat_plans[blk.pl.f_b_flat].flags & CF_FUSE)) CK(cuda_status(cudaStreamWaitEvent(s, m.ev_pack, 0)));
      CK(conv_bn(blk.b, blk.pl.f_b_flat, -1, s));
      if (fused_b) continue;   // the conv's second pass wrote blk.out and its ReLU bits
      if (ds_side) CK(cuda_status(cudaStreamWaitEvent(s, m.ev_pack, 0)));
      if (blk.has_ds) PROF(m, PC_BN_FWD, s, CK(run_bn_apply(m, B, blk.b.gout, blk.b.y, blk.b.bn, nullptr, blk.ds.y, &blk.ds.bn, blk.out, 1,
                                                            blk.bits_out, def_b, update_running, s)));
      else PROF(m, PC_BN_FWD, s, CK(run_bn_apply(m, B, blk.b.gout, blk.b.y, blk.b.bn, blk.in, nullptr, nullptr, blk.out, 1, blk.bits_out,
                                                 def_b, update_running, s)));
    }
  }
  avgpool_kernel<<<(B * 512 + 255) / 256, 256, 0, s>>>(m.blocks.back().out, m.feat, B, 512, m.blocks.back().b.gout); ++g_cilrs_launches;
  CKL();
  PROF(m, PC_HEADS, s, CK(heads_forward(m, B, speed, command, controls, pred_speed, keep_for_backward, dropout_p, seed, s, loss)));
  return OK;
}

static int heads_forward(Model& m, int B, const float* speed, const long long* command, float* controls, float* pred_speed,
                         int keep_for_backward, float dropout_p, unsigned long long seed, cudaStream_t s, const HeadsLossArgs* loss) {
  return heads_forward_run(heads_ctx(m), B, speed, command, controls, pred_speed, keep_for_backward, dropout_p, seed, loss, s);
}

// ---- backward building blocks ----
static int run_wgrad_old(Model& m, int idx, int slot, cudaStream_t s) {
  WgradParams wp = m.wold_plans[idx];
  wp.grad = m.grads + m.slots[slot].off;
  return launch_wgrad(&wp, s);
}
static int run_wgrad_flat(Model& m, int idx, cudaStream_t s) { return launch_wgrad_flat(&m.wflat_plans[idx], s); }
// flat dgrad with the fused ReLU mask + BatchNorm-backward reductions: bind workspace and dgamma / dbeta at launch time
static int launch_flat_bwd(Model& m, int idx, const BnRef* bn1, const BnRef* bn2, double count, cudaStream_t s) {
  FlatConvParams f = m.flat_plans[idx];
  if (f.flags & CF_BNBWD) {
    f.flags |= CF_DEFER;  // the bn_bwd_apply kernels that follow finalize (run_bn_bwd_apply, deferred)
    f.partials = bn1->bacc; f.counter = nullptr;
    f.dgamma1 = m.grads + m.slots[bn1->gamma].off; f.dbeta1 = m.grads + m.slots[bn1->beta].off;
    if (f.flags & CF_BNBWD2) { f.dgamma2 = m.grads + m.slots[bn2->gamma].off; f.dbeta2 = m.grads + m.slots[bn2->beta].off; }
    if (f.flags & CF_FUSE) {
      if (m.bar_bwd_next >= 64) return ERR_INVALID;
      f.grid_bar = m.bar_bwd + m.bar_bwd_next++;
      f.inv_count = 1.0 / count;
      f.gamma1 = m.params + m.slots[bn1->gamma].off;
      if (f.flags & CF_BNBWD2) f.gamma2 = m.params + m.slots[bn2->gamma].off;
    }
  }
  return launch_flat_conv(&f, s);
}

// standalone BatchNorm-backward reduce (where no flat dgrad produces the gradient): dz = g * (act > 0) written in place,
// bred = (sum dz, sum dz * xhat), dgamma / dbeta accumulated
static int run_bn_bwd_reduce(Model& m, int B, const PadGeom& g, const BnRef& bn, __nv_bfloat16* grad, const __nv_bfloat16* act,
                             const __nv_bfloat16* y, cudaStream_t s) {
  const long long nvec = pad_elems(B, g, bn.C) / 8;
  BnBwdReduceParams rp{};
  rp.g = grad; rp.act = act; rp.y = y; rp.mean = bn.vec + 2 * bn.C; rp.rstd = bn.vec + 3 * bn.C; rp.nvec = nvec; rp.C = bn.C;
  rp.partial = m.stat_acc; rp.counter = m.counters; rp.bsum = bn.bred; rp.bdot = bn.bred + bn.C;
  rp.dgamma = m.grads + m.slots[bn.gamma].off; rp.dbeta = m.grads + m.slots[bn.beta].off;
  rp.dz_out = grad; rp.geom = g;
  ++g_cilrs_launches;
  static const int red_cap = env_cap("CILRS_EW_REDUCE_CAP", 2);
  return cuda_status(launch_pdl(bn_bwd_reduce_kernel<false>, dim3(capped_grid(ew_reduce_grid(nvec, bn.C), red_cap)), dim3(EW_THREADS), 0, s, rp));
}

// dy = gamma * rstd * (dz - bsum/n - xhat * bdot/n)   (frozen: gamma * rstd * dz); dz is already ReLU-masked
// acc_sum / acc_dot != nullptr: the reductions are still the raw sums a fused dgrad epilogue left there (deferred finalize)
static int run_bn_bwd_apply(Model& m, int B, const PadGeom& g, const BnRef& bn, const __nv_bfloat16* dz, const __nv_bfloat16* y,
                            double count, int frozen, __nv_bfloat16* dy, const double* acc_sum, const double* acc_dot, cudaStream_t s) {
  const long long nvec = pad_elems(B, g, bn.C) / 8;
  BnBwdApplyParams ap{};
  ap.g = dz; ap.act = nullptr; ap.y = y; ap.mean = bn.vec + 2 * bn.C; ap.rstd = bn.vec + 3 * bn.C;
  ap.gamma = m.params + m.slots[bn.gamma].off; ap.bsum = bn.bred; ap.bdot = bn.bred + bn.C; ap.inv_count = (float)(1.0 / count);
  ap.frozen = frozen; ap.nvec = nvec; ap.C = bn.C; ap.dy = dy; ap.dz = nullptr; ap.geom = g;
  if (acc_sum) {
    ap.defer.acc_sum = acc_sum; ap.defer.acc_dot = acc_dot; ap.defer.bred = bn.bred;
    ap.defer.dgamma = m.grads + m.slots[bn.gamma].off; ap.defer.dbeta = m.grads + m.slots[bn.beta].off;
  }
  ++g_cilrs_launches;
  static int cap_b = -1;   // CTAs per SM the backward apply grid is capped at (2 = the resident CTAs; 1 measured the same, 4 and 8 slower)
  if (cap_b < 0) { const char* e = getenv("CILRS_EW_BWD_CAP"); cap_b = e ? atoi(e) : 2; if (cap_b < 1) cap_b = 2; }
  return cuda_status(launch_pdl(bn_bwd_apply_kernel<false>, dim3(ew_grid(nvec, bn.C, 4, cap_b)), dim3(EW_THREADS), 0, s, ap));
}

static int heads_backward(Model& m, int B, const float* dcontrols, const float* dspeed, const float* speed,
                          const long long* command, float dropout_p, cudaStream_t s, cudaStream_t ws) {
  return heads_backward_run(heads_ctx(m), B, dcontrols, dspeed, speed, command, dropout_p, s, ws, m.ev_heads);
}

static __nv_bfloat16* debug_gradient_buffer(Model& m, int hi) { return (hi >= 0 && ((15 - hi) & 1)) ? m.g1 : m.g0; }

// dbg_hi / dbg_lo (test hook, cilrs_model_debug_backward): run only blocks dbg_hi..max(dbg_lo,0) (none if dbg_hi < 0) from the
// gradient the caller placed in m.g0, plus the stem if dbg_lo < 0; the heads are skipped
static int backward(Model& m, int B, int mode, int part, const float* dcontrols, const float* dspeed, const float* speed,
                    const long long* command, float dropout_p, cudaStream_t s, bool async_part = false, int dbg_hi = -2,
                    int dbg_lo = -2) {
  const bool dbg = dbg_hi != -2;
  if (mode == MODE_INFER) return ERR_INVALID;
  if (B != m.planB || mode != m.planMode) return ERR_INVALID;  // must follow a forward with keep_for_backward
  if (!m.grads) return ERR_INVALID;
  const int frozen = mode == MODE_FROZEN;
  if (part < -1 || part > 6) return ERR_INVALID;   // 5 / 6: the two halves of part 4 (stem: BatchNorm backward | weight gradient)
  CK(pack_join(m, s));
  if (dbg) {
    CK(cuda_status(cudaMemsetAsync(m.acc_bwd, 0, (size_t)m.acc_bwd_bytes, s)));
    // the plans bake the ping-pong buffers in: block 15 reads g0 and writes g1, block 14 reads g1, ...
    __nv_bfloat16* gin = debug_gradient_buffer(m, dbg_hi);
    if (dbg_hi >= 0) {
      Block& top = m.blocks[dbg_hi];
      PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_reduce(m, B, top.b.gout, top.b.bn, gin, top.out, top.b.y, s)));
      if (top.has_ds) {
        // the downsample BatchNorm is fed by the same masked gradient: its reductions (the regular path gets them from the
        // fused dgrad of the next block)
        BnBwdReduceParams rp{};
        const long long nvec = pad_elems(B, top.b.gout, top.ds.bn.C) / 8;
        rp.g = gin; rp.act = nullptr; rp.y = top.ds.y; rp.mean = top.ds.bn.vec + 2 * top.ds.bn.C; rp.rstd = top.ds.bn.vec + 3 * top.ds.bn.C;
        rp.nvec = nvec; rp.C = top.ds.bn.C; rp.partial = m.stat_acc; rp.counter = m.counters; rp.bsum = top.ds.bn.bred;
        rp.bdot = top.ds.bn.bred + top.ds.bn.C; rp.dgamma = m.grads + m.slots[top.ds.bn.gamma].off;
        rp.dbeta = m.grads + m.slots[top.ds.bn.beta].off; rp.dz_out = nullptr; rp.geom = top.b.gout;
        ++g_cilrs_launches;
        CK(cuda_status(launch_pdl(bn_bwd_reduce_kernel<false>, dim3(ew_reduce_grid(nvec, rp.C)), dim3(EW_THREADS), 0, s, rp)));
      }
    }
    m.bw_gcur = gin; m.bw_gnext = gin == m.g0 ? m.g1 : m.g0;
    m.bw_deferred = false;
    m.bw_fused_prev = false;
    m.bar_bwd_next = 0;
  } else if (part <= 0) {
    CK(cuda_status(cudaMemsetAsync(m.acc_bwd, 0, (size_t)m.acc_bwd_bytes, s)));  // accumulators of the deferred BN-backward finalize
    {
      cudaStream_t hws = (m.side != nullptr && !m.prof.on) ? m.side : s;
      PROF(m, PC_HEADS, s, CK(heads_backward(m, B, dcontrols, dspeed, speed, command, dropout_p, s, hws)));
    }
    // ---- trunk: gradient of the last block's output, then its ReLU mask + BN_b reductions ----
    Block& last = m.blocks.back();
    avgpool_bwd_kernel<<<(int)((pad_elems(B, last.b.gout, 512) / 8 + 255) / 256), 256, 0, s>>>(m.dfeat, m.dfeat2, m.g0, B, 512, last.b.gout); ++g_cilrs_launches;
    CKL();
    PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_reduce(m, B, last.b.gout, last.b.bn, m.g0, last.out, last.b.y, s)));
    m.bw_gcur = m.g0; m.bw_gnext = m.g1;
    m.bw_deferred = false;
    m.bw_fused_prev = false;
    m.bar_bwd_next = 0;
  }
  __nv_bfloat16*& gcur = m.bw_gcur;
  __nv_bfloat16*& gnext = m.bw_gnext;
  // blocks 15..13 = layer4, 12..7 = layer3, 6..3 = layer2, 2..0 = layer1
  static const int part_hi[4] = {15, 12, 6, 2}, part_lo[4] = {13, 7, 3, 0};
  const int b_hi = dbg ? dbg_hi : (part < 0 ? 15 : (part < 4 ? part_hi[part] : -1));
  const int b_lo = dbg ? (dbg_lo < 0 ? 0 : dbg_lo) : (part < 0 ? 0 : (part < 4 ? part_lo[part] : 0));
  // Weight gradients (and their split-K reduction) run on the side stream: nothing on the dgrad / BatchNorm chain depends on
  // them, and their CTAs fill the SMs the chain leaves idle (98-CTA layer3 grids, HBM-bound elementwise kernels).
  const bool use_side = m.side != nullptr && !m.prof.on;
  cudaStream_t ws = use_side ? m.side : s;
  auto slot_write = [&](int slot) -> int {   // main stream is about to overwrite dy slot `slot`
    if (use_side && m.pending[slot]) {
      CK(cuda_status(cudaStreamWaitEvent(s, m.ev_done[slot], 0)));
      m.pending[slot] = false;
    }
    return OK;
  };
  auto slot_ready = [&](int slot) -> int {   // dy slot `slot` is complete on the main stream: the side stream may read it
    if (use_side) {
      CK(cuda_status(cudaEventRecord(m.ev_ready[slot], s)));
      CK(cuda_status(cudaStreamWaitEvent(m.side, m.ev_ready[slot], 0)));
    }
    return OK;
  };
  auto slot_read_done = [&](int slot) -> int {  // the side-stream readers of `slot` have been enqueued
    if (use_side) {
      CK(cuda_status(cudaEventRecord(m.ev_done[slot], m.side)));
      m.pending[slot] = true;
    }
    return OK;
  };
  for (int bi = b_hi; bi >= b_lo; --bi) {
    Block& blk = m.blocks[bi];
    const PadGeom& go = blk.b.gout;
    const double cnt = (double)B * blk.b.oh * blk.b.ow;
    const int sb = bi & 1, sa = 2 + (bi & 1);
    const int blk_part = bi >= 13 ? 0 : (bi >= 7 ? 1 : (bi >= 3 ? 2 : 3));
    const int sd = 4 + (blk_part & 1);
    __nv_bfloat16* dyb = m.dyb[bi & 1];
    __nv_bfloat16* dya = m.dya[bi & 1];
    __nv_bfloat16* dyd = m.dyd[blk_part & 1];
    // gcur = dz of this block's output (ReLU-masked), with the reductions of bn_b (and bn_ds) already in their bred
    const int Cb = blk.b.bn.C;
    if (!m.bw_fused_prev) {
      CK(slot_write(sb));
      const double* bsum_acc = m.bw_deferred ? blk.b.bn.bacc : nullptr;  // sum dz | sum dz*y_b | sum dz*y_ds
      PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_apply(m, B, go, blk.b.bn, gcur, blk.b.y, cnt, frozen, dyb, bsum_acc, bsum_acc ? bsum_acc + Cb : nullptr, s)));
      CK(slot_ready(sb));
      if (blk.has_ds) {
        CK(slot_write(sd));
        PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_apply(m, B, go, blk.ds.bn, gcur, blk.ds.y, cnt, frozen, dyd, bsum_acc, bsum_acc ? bsum_acc + 2 * Cb : nullptr, s)));
        CK(slot_ready(sd));
      }
    }  // else: the fused dgrad of the block above (CF_FUSE) already wrote dyb (and dyd) and signalled the slots
    PROF(m, PC_WGRAD, ws, CK(run_wgrad_flat(m, blk.pl.w_b, ws)));                     // dW_b
    CK(slot_read_done(sb));
    const bool fused_b = (m.flat_plans[blk.pl.d_b].flags & CF_FUSE) != 0;             // dgrad_b's second pass writes dy_a itself
    if (fused_b) CK(slot_write(sa));
    PROF(m, PC_DGRAD, s, CK(launch_flat_bwd(m, blk.pl.d_b, &blk.a.bn, nullptr, cnt, s)));  // ga = dz_a (+ BN_a reductions)
    if (!fused_b) {
      CK(slot_write(sa));
      PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_apply(m, B, go, blk.a.bn, m.ga, blk.a.y, cnt, frozen, dya, blk.a.bn.bacc, blk.a.bn.bacc + blk.a.bn.C, s)));
    }
    CK(slot_ready(sa));
    if (blk.pl.w_a_flat >= 0) PROF(m, PC_WGRAD, ws, CK(run_wgrad_flat(m, blk.pl.w_a_flat, ws)));
    else PROF(m, PC_WGRAD, ws, CK(run_wgrad_old(m, blk.pl.w_a_old, blk.a.w, ws)));
    CK(slot_read_done(sa));
    if (blk.has_ds) {
      PROF(m, PC_WGRAD, ws, CK(run_wgrad_old(m, blk.pl.w_ds_old, blk.ds.w, ws)));
      CK(slot_read_done(sd));
    }
    if (blk.pl.d_a_flat >= 0) {
      Block* pb = bi > 0 ? &m.blocks[bi - 1] : nullptr;
      const bool fused_a = (m.flat_plans[blk.pl.d_a_flat].flags & CF_FUSE) != 0;      // ... and dy of the block below's conv_b / downsample BatchNorms
      const int pb_part = (bi - 1) >= 13 ? 0 : ((bi - 1) >= 7 ? 1 : ((bi - 1) >= 3 ? 2 : 3));
      const int sbp = (bi - 1) & 1, sdp = 4 + (pb_part & 1);
      if (fused_a) {
        CK(slot_write(sbp));
        if (pb->has_ds) CK(slot_write(sdp));
      }
      PROF(m, PC_DGRAD, s, CK(launch_flat_bwd(m, blk.pl.d_a_flat, pb ? &pb->b.bn : nullptr, (pb && pb->has_ds) ? &pb->ds.bn : nullptr, cnt, s)));
      if (fused_a) {
        CK(slot_ready(sbp));
        if (pb->has_ds) CK(slot_ready(sdp));
      }
      m.bw_deferred = true;
      m.bw_fused_prev = fused_a;
    } else {
      PROF(m, PC_DGRAD, s, CK(launch_conv_gemm_multi(&m.old_plans[blk.pl.d_a_old], 4, s)));  // the four output parities
      // the parity launches write the raw gradient of the previous block's output: mask + reduce it here
      Block& pb = m.blocks[bi - 1];
      PROF(m, PC_BN_BWD, s, CK(run_bn_bwd_reduce(m, B, pb.b.gout, pb.b.bn, gnext, pb.out, pb.b.y, s)));
      m.bw_deferred = false;
      m.bw_fused_prev = false;
    }
    __nv_bfloat16* t = gcur; gcur = gnext; gnext = t;
    // the layer group's last weight gradient is queued: permute its accumulator tiles into the OIHW gradients now, under the
    // data-gradient chain of the layers below (all four at the end of the backward they fought the stem's HBM-bound tail)
    if (!dbg && bi == part_lo[blk_part]) PROF(m, PC_WGRAD, ws, CK(launch_wgrad_reduce(&m.red_jobs[blk_part], m.grads, ws)));
  }
  if (dbg)
    for (int pt = 0; pt < 4; ++pt) PROF(m, PC_WGRAD, ws, CK(launch_wgrad_reduce(&m.red_jobs[pt], m.grads, ws)));
  static const bool join_early = getenv("CILRS_JOIN_EARLY") != nullptr;   // measurement aid: the round-1 order (join before the stem)
  auto join_side = [&]() -> int {
    if (use_side && !async_part) {
      CK(cuda_status(cudaEventRecord(m.ev_join, m.side)));
      CK(cuda_status(cudaStreamWaitEvent(s, m.ev_join, 0)));
      for (int i = 0; i < 6; ++i) m.pending[i] = false;
    }
    return OK;
  };
  if (join_early) CK(join_side());
  // ---- stem: max-pool backward + ReLU + BN backward, then wgrad ----
  const bool stem_ew = dbg ? dbg_lo < 0 : (part < 0 || part == 4 || part == 5);
  const bool stem_wg = dbg ? dbg_lo < 0 : (part < 0 || part == 4 || part == 6);
  if (stem_ew) {
    const BnRef& bn = m.stem.bn;
    // sums over the POOL outputs (every pooled gradient lands on exactly one conv1 pixel, whose raw value the forward kept in
    // pool_ysel): the ordinary reduce kernel on (g, pool_out as the ReLU mask, ysel) - 57 MB instead of a pass over conv1's output
    const long long pvec = pad_elems(B, kGeom0, 64) / 8;
    BnBwdReduceParams rp{};
    rp.g = gcur; rp.act = m.pool_out; rp.y = m.pool_ysel; rp.mean = bn.vec + 2 * 64; rp.rstd = bn.vec + 3 * 64; rp.nvec = pvec; rp.C = 64;
    rp.partial = m.stat_acc; rp.counter = m.counters; rp.bsum = bn.bred; rp.bdot = bn.bred + 64;
    rp.dgamma = m.grads + m.slots[bn.gamma].off; rp.dbeta = m.grads + m.slots[bn.beta].off; rp.geom = kGeom0;
    rp.dz_out = gcur;   // in place: g * [pooled activation > 0] - the only place where the stem's ReLU mask matters
    PROF(m, PC_BN_BWD, s, { ++g_cilrs_launches; CK(cuda_status(launch_pdl(bn_bwd_reduce_kernel<false>, dim3(ew_reduce_grid(pvec, 64)), dim3(EW_THREADS), 0, s, rp))); });
    // route through the arg-max + ReLU mask + BatchNorm backward in one pass (stem_pool.cuh)
    StemBwdApplyParams ap{};
    ap.g = gcur; ap.argmax = m.pool_arg; ap.y = m.stem.y; ap.mean = rp.mean; ap.rstd = rp.rstd; ap.gamma = m.params + m.slots[bn.gamma].off;
    ap.bsum = rp.bsum; ap.bdot = rp.bdot; ap.inv_count = (float)(1.0 / ((double)B * 4400.0));
    ap.frozen = frozen; ap.B = B; ap.H = 44; ap.W = 100; ap.OH = 22; ap.OW = 50; ap.OHp = kGeom0.Hp; ap.OWp = kGeom0.Wp; ap.C = 64; ap.dy = m.dy_stem;
    const long long nblk = (long long)B * 22 * 50 * 8;
    static const int stem_cap = env_cap("CILRS_EW_STEM_CAP", 4);
    PROF(m, PC_BN_BWD, s, { ++g_cilrs_launches; CK(cuda_status(launch_pdl(stem_bwd_apply_kernel, dim3(capped_grid(ew_grid(nblk, 64, 2, 2), stem_cap)), dim3(EW_THREADS), 0, s, ap))); });
  }
  if (stem_wg) PROF(m, PC_WGRAD, s, CK(run_wgrad_old(m, m.stem_wgrad, m.stem.w, s)));
  // join: everything after this call on the caller's stream (allreduce of the part, Adam) sees the finished gradients.
  // AFTER the stem's kernels, which read nothing the gradient stream writes: joined before them, the stem's reduce waited
  // 25 us for layer1's last weight gradient and its fold (CUPTI timeline).
  if (!join_early) CK(join_side());
  if (use_side && async_part) {
    CK(cuda_status(cudaEventRecord(m.ev_part, s)));
    CK(cuda_status(cudaStreamWaitEvent(m.side, m.ev_part, 0)));
  }
  return OK;
}

}  // namespace cilrs

using namespace cilrs;

extern "C" {

struct cilrs_model {
  Model m;
};

int cilrs_model_param_layout(long long* offsets, long long* sizes, int capacity, long long* total_floats, long long* buffer_floats,
                             int* num_bn) {
  Model m;
  m.maxB = 1;
  build_topology(m, 1);
  const int n = (int)m.slots.size();
  for (int i = 0; i < n && i < capacity; ++i) {
    if (offsets) offsets[i] = m.slots[i].off;
    if (sizes) sizes[i] = m.slots[i].size;
  }
  if (total_floats) *total_floats = m.param_floats;
  if (buffer_floats) *buffer_floats = m.buffer_floats;
  if (num_bn) *num_bn = m.num_bn;
  return n;
}

size_t cilrs_model_workspace_bytes(int max_batch) {
  if (max_batch < 1) return 0;
  Model m;
  m.maxB = max_batch;
  build_topology(m, max_batch);
  return (size_t)carve(m, nullptr);
}

int cilrs_model_create(cilrs_model** out, int max_batch, void* workspace, size_t workspace_bytes, void* stream) {
  if (!out || max_batch < 1 || !workspace) return ERR_INVALID;
  if (((uintptr_t)workspace) & 1023) return ERR_INVALID;
  cilrs_model* h = new (std::nothrow) cilrs_model();
  if (!h) return ERR_INVALID;
  h->m.maxB = max_batch;
  build_topology(h->m, max_batch);
  const long long need = carve(h->m, (char*)workspace);
  if ((size_t)need > workspace_bytes) {
    delete h;
    return ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  // constants: reduction counters, unit scale/shift for the inference max-pool, error flag
  float unit[128];
  for (int i = 0; i < 64; ++i) { unit[i] = 1.f; unit[64 + i] = 0.f; }
  std::vector<PackJob> jobs;
  {
    int blocks = 0;
    auto add = [&](ConvRef& c) {
      PackJob j;
      j.w_off = h->m.slots[c.w].off; j.wf = c.wf; j.wd = c.wd; j.cout = c.d.out_c; j.cin = c.d.in_c; j.kk = c.d.kh * c.d.kw;
      j.first_block = blocks;
      blocks += (c.d.out_c / 32) * (c.d.in_c / 32);
      jobs.push_back(j);
    };
    int bi = 0;
    for (auto& blk : h->m.blocks) {
      h->m.pack_block_first[bi++] = blocks;
      add(blk.a);
      add(blk.b);
      if (blk.has_ds) add(blk.ds);
    }
    h->m.pack_block_first[16] = blocks;
    h->m.pack_njobs = (int)jobs.size();
    h->m.pack_blocks = blocks;
  }
  // the padding pixels of every padded-flat tensor must read as zero from the first step on (kernels keep them zero)
  int st = cuda_status(cudaMemsetAsync(workspace, 0, (size_t)need, s));
  if (!st) st = cuda_status(cudaMemsetAsync(h->m.counters, 0, 64, s));
  if (!st) st = cuda_status(cudaMemcpyAsync(h->m.pack_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, s));
  if (!st) st = cuda_status(cudaMemsetAsync(h->m.err_flag, 0, 64, s));
  if (!st) st = cuda_status(cudaMemcpyAsync(h->m.unit_vec, unit, sizeof(unit), cudaMemcpyHostToDevice, s));
  if (!st) st = cuda_status(cudaStreamSynchronize(s));  // `unit` is a stack buffer; creation is not on the hot path
  if (st) {
    delete h;
    return st;
  }
  if (!getenv("CILRS_NO_SIDE_STREAM")) {
    // side stream for the weight-gradient kernels (cilrs_model_backward forks and joins it on the caller's stream)
    // lowest priority: the dgrad / BatchNorm chain on the caller's stream is the critical path and gets the SMs first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->m.side, cudaStreamNonBlocking, prio_lo) != cudaSuccess) h->m.side = nullptr;
    bool ok = h->m.side != nullptr;
    for (int i = 0; ok && i < 6; ++i)
      ok = cudaEventCreateWithFlags(&h->m.ev_ready[i], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->m.ev_done[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->m.ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->m.ev_pack_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->m.ev_pack, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->m.ev_heads, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->m.ev_part, cudaEventDisableTiming) == cudaSuccess;
    if (!ok && h->m.side) { cudaStreamDestroy(h->m.side); h->m.side = nullptr; }
  }
  *out = h;
  return OK;
}

void cilrs_model_destroy(cilrs_model* h) {
  if (!h) return;
  if (h->m.side) {
    cudaStreamSynchronize(h->m.side);
    for (int i = 0; i < 6; ++i) {
      if (h->m.ev_ready[i]) cudaEventDestroy(h->m.ev_ready[i]);
      if (h->m.ev_done[i]) cudaEventDestroy(h->m.ev_done[i]);
    }
    if (h->m.ev_join) cudaEventDestroy(h->m.ev_join);
    if (h->m.ev_pack_fork) cudaEventDestroy(h->m.ev_pack_fork);
    if (h->m.ev_pack) cudaEventDestroy(h->m.ev_pack);
    if (h->m.ev_heads) cudaEventDestroy(h->m.ev_heads);
    if (h->m.ev_part) cudaEventDestroy(h->m.ev_part);
    cudaStreamDestroy(h->m.side);
  }
  delete h;
}

int cilrs_model_bind(cilrs_model* h, float* params, float* grads, float* buffers, long long* num_batches_tracked) {
  if (!h || !params || !buffers) return ERR_INVALID;
  if ((((uintptr_t)params) | ((uintptr_t)grads) | ((uintptr_t)buffers)) & 15) return ERR_INVALID;
  Model& m = h->m;
  m.params = params; m.grads = grads; m.buffers = buffers; m.nbt = num_batches_tracked;
  return OK;
}

int cilrs_model_refresh(cilrs_model* h, int what, void* stream) {
  if (!h) return ERR_INVALID;
  return refresh(h->m, what, (cudaStream_t)stream);
}

// what = 1 of cilrs_model_refresh, off the critical path: the stem's operand is packed on `stream`, the 36 trunk convolutions'
// on the model's gradient stream (forked behind everything queued on `stream`); the next cilrs_model_forward* on `stream`
// waits for them only after its stem convolution + pooling (59 us of packing under 100 us of stem at batch 128).
int cilrs_model_refresh_async(cilrs_model* h, void* stream) {
  if (!h || !h->m.params) return ERR_INVALID;
  Model& m = h->m;
  cudaStream_t s = (cudaStream_t)stream;
  if (!m.side || m.prof.on) return refresh(m, 1, s);
  CK(pack_join(m, s));
  // the stem's 64-CTA repack goes first: queued behind the trunk's 2456 CTAs it waited 40 us for a free slot (CUPTI timeline),
  // and K0 + the whole forward behind it
  CK(cilrs_stem_pack_weight(m.params + m.slots[m.stem.w].off, m.stem.wf, s));
  CK(cuda_status(cudaEventRecord(m.ev_pack_fork, s)));
  CK(cuda_status(cudaStreamWaitEvent(m.side, m.ev_pack_fork, 0)));
  CK(launch_pack_all(m.params, m.pack_jobs, m.pack_njobs, m.pack_blocks, m.side));
  CK(cuda_status(cudaEventRecord(m.ev_pack, m.side)));
  m.pack_pending = true;
  return OK;
}

// repack the bf16 operands of the convolutions whose gradients backward part `part` completes (0 = layer4, 1 = layer3,
// 2 = layer2, 3 = layer1, 4 = stem): lets the optimizer + repack of a finished part run under the rest of the backward
int cilrs_model_refresh_part(cilrs_model* h, int part, void* stream) {
  if (!h || part < 0 || part > 4 || !h->m.params) return ERR_INVALID;
  Model& m = h->m;
  cudaStream_t s = (cudaStream_t)stream;
  if (part == 4) return cilrs_stem_pack_weight(m.params + m.slots[m.stem.w].off, m.stem.wf, s);
  static const int lo[4] = {13, 7, 3, 0}, hi[4] = {16, 13, 7, 3};
  return launch_pack_range(m.params, m.pack_jobs, m.pack_njobs, m.pack_block_first[lo[part]], m.pack_block_first[hi[part]], s);
}

int cilrs_model_forward(cilrs_model* h, int batch, int mode, const float* image_nchw, const void* image_s2d, const float* speed,
                        const long long* command, float* controls, float* pred_speed, int update_running_stats,
                        int keep_for_backward, float dropout_p, unsigned long long seed, void* stream) {
  if (!h || !speed || !command || !controls || !pred_speed) return ERR_INVALID;
  if (mode < 0 || mode > 2) return ERR_INVALID;
  if (mode == MODE_INFER && keep_for_backward) return ERR_INVALID;
  return forward(h->m, batch, mode, image_nchw, image_s2d, speed, command, controls, pred_speed, update_running_stats,
                 keep_for_backward, dropout_p, seed, (cudaStream_t)stream);
}

// forward + the training loss fused into the heads kernel (the last cluster to finish reduces over the batch): one launch
// less on the step's critical path. The speed head's target is the speed INPUT (notebook/notebook.ipynb:550).
int cilrs_model_forward_loss(cilrs_model* h, int batch, int mode, const float* image_nchw, const void* image_s2d, const float* speed,
                             const long long* command, float* controls, float* pred_speed, int update_running_stats,
                             float dropout_p, unsigned long long seed, const float* targets, int loss_mode, float w_steer,
                             float w_throttle, float w_brake, float w_speed, float grad_scale, float* out6, float* dcontrols,
                             float* dspeed, void* stream) {
  if (!h || !speed || !command || !controls || !pred_speed || !targets || !out6) return ERR_INVALID;
  if (mode != MODE_TRAIN && mode != MODE_FROZEN) return ERR_INVALID;
  if (loss_mode != 0 && loss_mode != 1) return ERR_INVALID;
  HeadsLossArgs la;
  la.targets = targets; la.speed_target = speed; la.mode = loss_mode; la.w_steer = w_steer; la.w_throttle = w_throttle;
  la.w_brake = w_brake; la.w_speed = w_speed; la.grad_scale = grad_scale; la.out6 = out6; la.dcontrols = dcontrols; la.dspeed = dspeed;
  return forward(h->m, batch, mode, image_nchw, image_s2d, speed, command, controls, pred_speed, update_running_stats, 1, dropout_p,
                 seed, (cudaStream_t)stream, &la);
}

int cilrs_model_backward(cilrs_model* h, int batch, int mode, int part, const float* dcontrols, const float* dspeed,
                         const float* speed, const long long* command, float dropout_p, void* stream) {
  if (!h || !dcontrols || !dspeed || !speed || !command) return ERR_INVALID;
  return backward(h->m, batch, mode, part, dcontrols, dspeed, speed, command, dropout_p, (cudaStream_t)stream);
}

int cilrs_model_backward_part_async(cilrs_model* h, int batch, int mode, int part, const float* dcontrols, const float* dspeed,
                                    const float* speed, const long long* command, float dropout_p, void* stream) {
  if (!h || !dcontrols || !dspeed || !speed || !command || part < 0) return ERR_INVALID;
  return backward(h->m, batch, mode, part, dcontrols, dspeed, speed, command, dropout_p, (cudaStream_t)stream, true);
}
void* cilrs_model_gradient_stream(cilrs_model* h) { return (h && h->m.side && !h->m.prof.on) ? (void*)h->m.side : nullptr; }
int cilrs_model_backward_join(cilrs_model* h, void* stream) {
  if (!h) return ERR_INVALID;
  Model& m = h->m;
  if (m.side && !m.prof.on) {
    CK(cuda_status(cudaEventRecord(m.ev_join, m.side)));
    CK(cuda_status(cudaStreamWaitEvent((cudaStream_t)stream, m.ev_join, 0)));
    for (int i = 0; i < 6; ++i) m.pending[i] = false;
  }
  return OK;
}

int cilrs_model_backward_part_first_tensor(int part) {
  static const int first_block[5] = {13, 7, 3, 0, -1};
  if (part < 0 || part > 4) return -1;
  if (part == 4) return 0;
  Model m;
  m.maxB = 1;
  build_topology(m, 1);
  return m.blocks[first_block[part]].a.w;
}

// per-class kernel timing: enable, run forward/backward (not inside a timed region), then collect (synchronises)
int cilrs_model_profile(cilrs_model* h, int enable) {
  if (!h) return ERR_INVALID;
  h->m.prof.on = enable != 0;
  h->m.prof.used = 0;
  h->m.prof.recs.clear();
  return OK;
}
// out_ms[7], out_launches[7]: fprop, dgrad, wgrad, bn forward (+pool), bn backward, heads, other
int cilrs_model_profile_collect(cilrs_model* h, float* out_ms, int* out_launches) {
  if (!h || !out_ms || !out_launches) return ERR_INVALID;
  for (int i = 0; i < PC_COUNT; ++i) { out_ms[i] = 0.f; out_launches[i] = 0; }
  for (auto& r : h->m.prof.recs) {
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return cuda_status(e);
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, r.a, r.b);
    if (e != cudaSuccess) return cuda_status(e);
    out_ms[r.cls] += ms;
    out_launches[r.cls] += 1;
  }
  h->m.prof.used = 0;
  h->m.prof.recs.clear();
  return OK;
}

// heads only (test hooks + building block): features [B,512] f32 in, like the trunk would leave them
int cilrs_model_heads_forward(cilrs_model* h, int batch, const float* feat, const float* speed, const long long* command,
                              float* controls, float* pred_speed, int keep_for_backward, float dropout_p, unsigned long long seed,
                              void* stream) {
  if (!h || !feat || !speed || !command || !controls || !pred_speed) return ERR_INVALID;
  Model& m = h->m;
  if (batch < 1 || batch > m.maxB || !m.params) return ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  if (feat != m.feat) CK(cuda_status(cudaMemcpyAsync(m.feat, feat, (size_t)batch * 512 * 4, cudaMemcpyDeviceToDevice, s)));
  return heads_forward(m, batch, speed, command, controls, pred_speed, keep_for_backward, dropout_p, seed, s);
}
int cilrs_model_heads_backward(cilrs_model* h, int batch, const float* dcontrols, const float* dspeed, const float* speed,
                               const long long* command, float dropout_p, float* dfeat_out, void* stream) {
  if (!h || !dcontrols || !dspeed || !speed || !command) return ERR_INVALID;
  Model& m = h->m;
  if (batch < 1 || batch > m.maxB || !m.params || !m.grads) return ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  CK(heads_backward(m, batch, dcontrols, dspeed, speed, command, dropout_p, s, s));
  if (dfeat_out) {
    add2_kernel<<<(batch * 512 + 255) / 256, 256, 0, s>>>(m.dfeat, m.dfeat2, dfeat_out, (long long)batch * 512); ++g_cilrs_launches;
    CKL();
  }
  return OK;
}

// debug / test hook: device pointer and NHWC dims of an intermediate activation of the last forward
//   which: 0 = max-pool output, 1..16 = BasicBlock outputs, 17 = stem conv raw output;
//   dims = {H, W, C, Hp, Wp}: the tensor is [batch, Hp, Wp, C] with the real pixels in [:, :H, :W] (padded-flat layout)
void* cilrs_model_debug_activation(cilrs_model* h, int which, int* dims) {
  if (!h || !dims) return nullptr;
  Model& m = h->m;
  if (which == 0) { dims[0] = 22; dims[1] = 50; dims[2] = 64; dims[3] = kGeom0.Hp; dims[4] = kGeom0.Wp; return m.pool_out; }
  if (which >= 1 && which <= (int)m.blocks.size()) {
    Block& b = m.blocks[which - 1];
    dims[0] = b.b.oh; dims[1] = b.b.ow; dims[2] = b.b.d.out_c; dims[3] = b.b.gout.Hp; dims[4] = b.b.gout.Wp;
    return b.out;
  }
  if (which == 17) { dims[0] = 44; dims[1] = 100; dims[2] = 64; dims[3] = 44; dims[4] = 100; return m.stem.y; }
  return nullptr;
}

// test hook: backward of blocks hi..max(lo,0) only (hi < 0: none), plus the stem when lo < 0, from a given gradient.
// g_out: bf16 padded-flat gradient w.r.t. the OUTPUT of block hi (before that block's final ReLU mask is applied), or - stem only -
// w.r.t. the max-pool output [batch,23,51,64]; its padding pixels must be zero. Must follow a forward(keep_for_backward) of the
// same batch / mode. Parameter gradients are accumulated into the bound arena as usual; the gradient w.r.t. the input of block
// max(lo,0) is left in cilrs_model_debug_gradient() (same padded-flat geometry as that input; masked by the previous block's ReLU
// when lo > 0, exactly as the full backward does).
int cilrs_model_debug_backward(cilrs_model* h, int batch, int mode, int hi, int lo, const void* g_out, void* stream) {
  if (!h || !g_out || hi > 15 || lo > hi + 1 || lo < -1 || hi < -1) return ERR_INVALID;
  Model& m = h->m;
  if (hi < 0 && lo >= 0) return ERR_INVALID;
  const PadGeom g = hi >= 0 ? m.blocks[hi].b.gout : kGeom0;
  const int C = hi >= 0 ? m.blocks[hi].b.d.out_c : 64;
  CK(cuda_status(cudaMemcpyAsync(debug_gradient_buffer(m, hi), g_out, (size_t)pad_elems(batch, g, C) * 2, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream)));
  return backward(m, batch, mode, -1, nullptr, nullptr, nullptr, nullptr, 0.f, (cudaStream_t)stream, false, hi, lo);
}
void* cilrs_model_debug_gradient(cilrs_model* h) { return h ? (void*)h->m.bw_gcur : nullptr; }
// test hook: fp32 [batch, width] head activations kept by the last forward(keep_for_backward), post-ReLU and post-dropout.
// which: 0 = speed_encoder.0 (128), 1 = speed_encoder.3 (128), 2 = branch.0 (256), 3 = branch.3 (256), 4 = speed_predictor.0
// (256), 5 = speed_predictor.3 (256); *width receives the row length
float* cilrs_model_debug_heads_saved(cilrs_model* h, int which, int* width) {
  if (!h || which < 0 || which > 5) return nullptr;
  HeadsSaved& s = h->m.hs;
  float* ptr[6] = {s.s1, s.sfeat, s.b1, s.b2, s.p1, s.p2};
  if (width) *width = which < 2 ? 128 : 256;
  return ptr[which];
}

// dropout under a captured CUDA graph: the mask seed of every forward becomes seed + golden * (*counter + 1), with `counter` a
// device int64 that changes between replays (FusedTrainer passes the optimizer's device step counter). NULL switches it off.
int cilrs_model_set_dropout_counter(cilrs_model* h, const long long* counter_dev) {
  if (!h) return ERR_INVALID;
  h->m.drop_counter = counter_dev;
  return OK;
}

// forget the cached plans (tensor maps, tile shapes, fusion decisions): the next forward rebuilds them
int cilrs_model_invalidate_plans(cilrs_model* h) {
  if (!h) return ERR_INVALID;
  h->m.planB = 0; h->m.planMode = -1;
  return OK;
}

void* cilrs_model_input_s2d(cilrs_model* h) { return h ? (void*)h->m.x_s2d : nullptr; }
int* cilrs_model_error_flag(cilrs_model* h) { return h ? h->m.err_flag : nullptr; }


// ------------------------------------------------------------------------------------------------
// single-kernel entry points of the BatchNorm / pooling family (used by the unit tests; the network plan above
// launches the same kernels). `vec` is [4][C] fp32: scale, shift, mean, rstd.
// ------------------------------------------------------------------------------------------------
int cilrs_bn_finalize(const float* partials, int tiles, int C, double count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                      int training, int update_running, float* vec, void* stream) {
  if (!gamma || !beta || !running_mean || !running_var || !vec || C < 1 || (training && (!partials || tiles < 1))) return ERR_INVALID;
  BnVectors v{vec, vec + C, vec + 2 * C, vec + 3 * C};
  bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partials, tiles, C, count, gamma, beta, running_mean,
                                                                       running_var, num_batches_tracked, momentum, eps, training,
                                                                       update_running, v); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

// pad_h, pad_w > 0: the tensors are padded-flat [batch, pad_h + 1, pad_w + 1, C] (elems counts the padding pixels too)
static int abi_geom(int pad_h, int pad_w, long long elems, int C, PadGeom* g) {
  if (pad_h <= 0 && pad_w <= 0) { *g = kDense; return OK; }
  if (pad_h <= 0 || pad_w <= 0) return ERR_INVALID;
  *g = PadGeom{pad_h, pad_w, pad_h + 1, pad_w + 1};
  return (elems % ((long long)g->Hp * g->Wp * C)) ? ERR_INVALID : OK;
}

int cilrs_bn_apply(const void* x, const float* vec, const void* residual, const void* x2, const float* vec2, void* out,
                   long long elems, int C, int relu, int pad_h, int pad_w, uint8_t* relu_bits, void* stream) {
  if (!x || !vec || !out || C < 64 || C % 64 || elems % C) return ERR_INVALID;
  PadGeom g;
  CK(abi_geom(pad_h, pad_w, elems, C, &g));
  const long long nvec = elems / 8;
  bn_apply_kernel<<<ew_grid(nvec, C), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, vec, vec + C, (const __nv_bfloat16*)residual, (const __nv_bfloat16*)x2, vec2, vec2 ? vec2 + C : nullptr,
      (__nv_bfloat16*)out, nvec, C, relu, g, relu_bits, BnDefer{}); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

// padded_out != 0: out (and ysel) are padded-flat [batch, OH + 1, OW + 1, C] (their padding pixels are not written); argmax stays
// dense. ysel (optional, needs argmax): the raw y at the arg-max of every window, for cilrs_stem_bn_backward.
int cilrs_bn_relu_maxpool_sel(const void* y, const float* vec, void* out, uint8_t* argmax, void* ysel, int batch, int H, int W, int C,
                              int padded_out, void* stream) {
  if (!y || !vec || !out || C % 64 || batch < 1 || H < 1 || W < 1 || (ysel && !argmax)) return ERR_INVALID;
  if ((long long)batch * (H + 2) * (W + 2) * C >= (1LL << 31)) return ERR_UNSUPPORTED;   // the kernel indexes with 32 bits
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long nvec = (long long)batch * OH * OW * C / 8;
  bn_relu_maxpool_sel_kernel<<<ew_grid(nvec, C), EW_THREADS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y, vec, vec + C, (__nv_bfloat16*)out, argmax, (__nv_bfloat16*)ysel, batch, H, W, C, OH, OW,
      padded_out ? OH + 1 : OH, padded_out ? OW + 1 : OW); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
int cilrs_bn_relu_maxpool(const void* y, const float* vec, void* out, uint8_t* argmax, int batch, int H, int W, int C, int padded_out,
                          void* stream) {
  return cilrs_bn_relu_maxpool_sel(y, vec, out, argmax, nullptr, batch, H, W, C, padded_out, stream);
}

// The stem's max-pool + ReLU + BatchNorm backward as the training plan runs it (stem_pool.cuh): sums over the pool outputs
// (g, act = the pooled activations, ysel from cilrs_bn_relu_maxpool_sel), then one routing + apply pass. H and W even.
// g / act / ysel: [batch, H/2, W/2, C], padded-flat [batch, H/2 + 1, W/2 + 1, C] when padded != 0; y, dy: dense [batch, H, W, C].
// workspace / counter as for cilrs_bn_backward.
int cilrs_stem_bn_backward(void* g, const void* act, const void* ysel, const uint8_t* argmax, const void* y, const float* vec,
                           const float* gamma, int batch, int H, int W, int C, int padded, int frozen, void* dy, float* dgamma,
                           float* dbeta, float* workspace, unsigned int* counter, void* stream) {
  if (!g || !act || !ysel || !argmax || !y || !vec || !gamma || !dy || !workspace || !counter) return ERR_INVALID;
  if (C % 64 || batch < 1 || H < 2 || W < 2 || (H & 1) || (W & 1)) return ERR_INVALID;
  if ((long long)batch * (H + 2) * (W + 2) * C >= (1LL << 31)) return ERR_UNSUPPORTED;   // the kernel indexes with 32 bits
  cudaStream_t s = (cudaStream_t)stream;
  const int OH = H / 2, OW = W / 2;
  const PadGeom geom = padded ? PadGeom{OH, OW, OH + 1, OW + 1} : kDense;
  const long long pvec = (long long)batch * (padded ? (OH + 1) * (OW + 1) : OH * OW) * C / 8;
  float* bred = workspace + (size_t)EW_MAX_BLOCKS * 2 * C;
  BnBwdReduceParams rp{};
  rp.g = (const __nv_bfloat16*)g; rp.act = (const __nv_bfloat16*)act; rp.y = (const __nv_bfloat16*)ysel;
  rp.mean = vec + 2 * C; rp.rstd = vec + 3 * C; rp.nvec = pvec; rp.C = C; rp.partial = (double*)workspace; rp.counter = counter;
  rp.bsum = bred; rp.bdot = bred + C; rp.dgamma = dgamma; rp.dbeta = dbeta; rp.geom = geom;
  rp.dz_out = (__nv_bfloat16*)g;   // g becomes g * [act > 0] in place
  bn_bwd_reduce_kernel<false><<<ew_reduce_grid(pvec, C), EW_THREADS, 0, s>>>(rp); ++g_cilrs_launches;
  CKL();
  StemBwdApplyParams ap{};
  ap.g = rp.g; ap.argmax = argmax; ap.y = (const __nv_bfloat16*)y; ap.mean = rp.mean; ap.rstd = rp.rstd; ap.gamma = gamma;
  ap.bsum = rp.bsum; ap.bdot = rp.bdot; ap.inv_count = (float)(1.0 / ((double)batch * H * W));
  ap.frozen = frozen; ap.B = batch; ap.H = H; ap.W = W; ap.OH = OH; ap.OW = OW; ap.OHp = padded ? OH + 1 : OH; ap.OWp = padded ? OW + 1 : OW;
  ap.C = C; ap.dy = (__nv_bfloat16*)dy;
  const long long nblk = (long long)batch * OH * OW * (C / 8);
  stem_bwd_apply_kernel<<<ew_grid(nblk, C, 2, 2), EW_THREADS, 0, s>>>(ap); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

// BN (+ optional ReLU mask from `act`) backward. workspace: cilrs_bn_backward_workspace_floats(C) floats whose first 2*C are
// ZERO on entry (global accumulators; left zero) + one zeroed uint32 counter.
// stem variant (argmax != NULL): g is the pooled gradient [B,(H+1)/2,(W+1)/2,C] routed through the 3x3/2 max-pool; with
// pad_h > 0 that pooled gradient is padded-flat [B,(H+1)/2+1,(W+1)/2+1,C].
// regular variant with pad_h, pad_w > 0: g / act / y / dy / dz are padded-flat [batch, pad_h+1, pad_w+1, C].
int cilrs_bn_backward(const void* g, const void* act, const void* y, const float* vec, const float* gamma, long long elems, int C,
                      double count, int frozen, void* dy, void* dz, float* dgamma, float* dbeta, float* workspace,
                      unsigned int* counter, const uint8_t* argmax, int H, int W, int pad_h, int pad_w, void* stream) {
  if (!g || !y || !vec || !gamma || !dy || !workspace || !counter || C % 64 || elems % C) return ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PadGeom geom = kDense;
  if (!argmax) CK(abi_geom(pad_h, pad_w, elems, C, &geom));
  const long long nvec = elems / 8;
  const int grid = ew_grid(nvec, C);
  const int rgrid = ew_reduce_grid(nvec, C);
  float* bred = workspace + (size_t)EW_MAX_BLOCKS * 2 * C;
  BnBwdReduceParams rp{};
  rp.g = (const __nv_bfloat16*)g; rp.act = (const __nv_bfloat16*)act; rp.y = (const __nv_bfloat16*)y;
  rp.mean = vec + 2 * C; rp.rstd = vec + 3 * C; rp.nvec = nvec; rp.C = C; rp.partial = (double*)workspace; rp.counter = counter;
  rp.bsum = bred; rp.bdot = bred + C; rp.dgamma = dgamma; rp.dbeta = dbeta; rp.geom = geom;
  BnBwdApplyParams ap{};
  ap.g = rp.g; ap.act = rp.act; ap.y = rp.y; ap.mean = rp.mean; ap.rstd = rp.rstd; ap.gamma = gamma; ap.bsum = rp.bsum; ap.bdot = rp.bdot;
  ap.inv_count = (float)(1.0 / count); ap.frozen = frozen; ap.nvec = nvec; ap.C = C; ap.dy = (__nv_bfloat16*)dy; ap.dz = (__nv_bfloat16*)dz;
  ap.geom = geom;
  if (argmax) {
    rp.argmax = argmax; rp.scale = vec; rp.shift = vec + C; rp.H = H; rp.W = W; rp.OH = (H + 1) / 2; rp.OW = (W + 1) / 2;
    rp.OHp = pad_h > 0 ? rp.OH + 1 : rp.OH; rp.OWp = pad_h > 0 ? rp.OW + 1 : rp.OW;
    rp.dz_out = (__nv_bfloat16*)dy;  // routed + masked gradient; the apply pass turns it into dy in place
    bn_bwd_reduce_kernel<true><<<ew_grid(nvec, C, 8), EW_THREADS, 0, s>>>(rp); ++g_cilrs_launches;
    CKL();
    ap.g = (const __nv_bfloat16*)dy; ap.act = nullptr;
    bn_bwd_apply_kernel<false><<<grid, EW_THREADS, 0, s>>>(ap); ++g_cilrs_launches;
  } else {
    bn_bwd_reduce_kernel<false><<<rgrid, EW_THREADS, 0, s>>>(rp); ++g_cilrs_launches;
    CKL();
    bn_bwd_apply_kernel<false><<<grid, EW_THREADS, 0, s>>>(ap); ++g_cilrs_launches;
  }
  return cuda_status(cudaGetLastError());
}
size_t cilrs_bn_backward_workspace_floats(int C) { return (size_t)EW_MAX_BLOCKS * 2 * C + 2 * (size_t)C; }

int cilrs_loss(const float* controls, const float* pred_speed, const float* targets, const float* speed_target, int batch, int mode,
               float w_steer, float w_throttle, float w_brake, float w_speed, float grad_scale, float* out6, float* dcontrols,
               float* dspeed, void* stream) {
  if (!controls || !pred_speed || !targets || !speed_target || !out6 || batch < 1) return ERR_INVALID;
  if (mode != 0 && mode != 1) return ERR_INVALID;
  LossParams p;
  p.controls = controls; p.pred_speed = pred_speed; p.targets = targets; p.speed_target = speed_target; p.batch = batch; p.mode = mode;
  p.w_steer = w_steer; p.w_throttle = w_throttle; p.w_brake = w_brake; p.w_speed = w_speed; p.grad_scale = grad_scale;
  p.out = out6; p.dcontrols = dcontrols; p.dspeed = dspeed;
  loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

int cilrs_validate_accumulate(const float* controls, const float* pred_speed, const float* targets, const float* speed_target,
                              const long long* command, int batch, int mode, float w_steer, float w_throttle, float w_brake,
                              float w_speed, double* acc16, void* stream) {
  if (!controls || !pred_speed || !targets || !speed_target || !command || !acc16 || batch < 1) return ERR_INVALID;
  if (mode != 0 && mode != 1) return ERR_INVALID;
  ValidateParams p;
  p.controls = controls; p.pred_speed = pred_speed; p.targets = targets; p.speed_target = speed_target; p.command = command;
  p.batch = batch; p.mode = mode; p.w_steer = w_steer; p.w_throttle = w_throttle; p.w_brake = w_brake; p.w_speed = w_speed; p.acc = acc16;
  validate_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

static int adam_launch(AdamParams& a, long long* step_dev, cudaStream_t s, bool advance = true) {
  if (step_dev && advance) {
    step_increment_kernel<<<1, 1, 0, s>>>(step_dev); ++g_cilrs_launches;
    int st = cuda_status(cudaGetLastError());
    if (st) return st;
  }
  long long blocks = (a.n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<(int)blocks, 256, 0, s>>>(a); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

int cilrs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, long long step, long long* step_dev, float grad_scale, const float* grad_scale_dev,
                    void* stream) {
  if (!p || !g || !m || !v || n < 0 || (n & 3) || (step < 1 && !step_dev)) return ERR_INVALID;
  if ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) return ERR_INVALID;
  if (n == 0) return OK;
  AdamParams a{};
  a.p = p; a.g = const_cast<float*>(g); a.m = m; a.v = v; a.n = n; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_correction1 = step >= 1 ? (float)(1.0 - pow((double)beta1, (double)step)) : 1.f;
  a.bias_correction2_sqrt = step >= 1 ? (float)sqrt(1.0 - pow((double)beta2, (double)step)) : 1.f;
  a.grad_scale = grad_scale; a.grad_scale_dev = grad_scale_dev; a.step_dev = step_dev;
  return adam_launch(a, step_dev, (cudaStream_t)stream);
}

// The CUDA-graph form: every hyper-parameter lives in device memory (hyper_dev float[8] = lr, beta1, beta2, eps, weight_decay,
// grad_scale, -, -; step_dev int64 incremented on the stream first), so one captured graph follows an LR schedule and resumes
// from a loaded optimizer state. g_bf16 (optional) replaces g as the gradient source (all-reduced bf16 exchange buffer);
// zero_grad also clears the fp32 arena g (the next step's optimizer.zero_grad()).
int cilrs_adam_step_ex(float* p, float* g, const void* g_bf16, float* m, float* v, long long n, const float* hyper_dev,
                       long long* step_dev, const float* grad_scale_dev, int zero_grad, void* stream) {
  if (!p || !m || !v || n < 0 || (n & 3) || !hyper_dev || !step_dev) return ERR_INVALID;
  if (!g && (!g_bf16 || zero_grad)) return ERR_INVALID;
  if ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v) | ((uintptr_t)hyper_dev)) & 15) return ERR_INVALID;
  if (((uintptr_t)g_bf16) & 7) return ERR_INVALID;
  if (n == 0) return OK;
  AdamParams a{};
  a.p = p; a.g = g; a.g16 = (const __nv_bfloat16*)g_bf16; a.m = m; a.v = v; a.n = n;
  a.bias_correction1 = 1.f; a.bias_correction2_sqrt = 1.f; a.grad_scale = 1.f;
  a.grad_scale_dev = grad_scale_dev; a.step_dev = step_dev; a.hyper_dev = hyper_dev; a.zero_grad = zero_grad & 1;
  return adam_launch(a, step_dev, (cudaStream_t)stream, !(zero_grad & 2));
}

static int sumsq_launch(const float* g, const void* g16, long long n, double* partial_ws, unsigned int* counter_ws, float max_norm,
                        float* out2, void* stream) {
  if ((!g && !g16) || !partial_ws || !counter_ws || !out2 || n < 0 || (n & 3)) return ERR_INVALID;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)g16, n, partial_ws, counter_ws, out2, max_norm); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}
int cilrs_grad_sumsq(const float* g, long long n, double* partial_ws /* >= 1024 doubles */, unsigned int* counter_ws /* zeroed */,
                     float max_norm, float* out2, void* stream) {
  return sumsq_launch(g, nullptr, n, partial_ws, counter_ws, max_norm, out2, stream);
}
int cilrs_grad_sumsq_bf16(const void* g_bf16, long long n, double* partial_ws, unsigned int* counter_ws, float max_norm, float* out2,
                          void* stream) {
  return sumsq_launch(nullptr, g_bf16, n, partial_ws, counter_ws, max_norm, out2, stream);
}

int cilrs_grad_to_bf16(float* g, void* out_bf16, long long n, int zero_source, void* stream) {
  if (!g || !out_bf16 || n < 0 || (n & 3) || (((uintptr_t)g) & 15) || (((uintptr_t)out_bf16) & 7)) return ERR_INVALID;
  if (n == 0) return OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  grad_to_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, (__nv_bfloat16*)out_bf16, n, zero_source); ++g_cilrs_launches;
  return cuda_status(cudaGetLastError());
}

}  // extern "C"
