// rfx_group.inl — row-sharded multi-GPU groups behind the C ABI (included at the end of rfx_api.cu: it needs the private
// definitions of rfx_ctx / rfx_ssgi_chain).
//
// One process per GPU.  A group owns
//   * an NCCL communicator (bootstrap from a 128-byte unique id the host shares by any means) used for ONE tiny collective per
//     frame — an all-gather of each rank's kernel time, which is also the frame barrier — and for the exchange of the CUDA IPC
//     handles at attach time;
//   * peer mappings (cudaIpcOpenMemHandle) of every rank's double-buffered history planes (`composed`, `dn`), so that K1 and K2
//     of the next frame read the rows another rank owns in place over NVLink instead of receiving replicated planes;
//   * the band table (one contiguous row band per rank) and its cost-driven rebalancing.
// Bounded stencils (K2's 5x5 window, the Poisson taps, K4) are recomputed locally on widened row ranges (rfx_shard_ranges),
// so the frame needs no per-pass exchange and its result is bit-identical to the single-GPU chain.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  // a process that already loaded NCCL (e.g. torch's bundled copy) resolves the soname to that copy; otherwise the system library
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) return &api;
#define RFX_SYM(n) *(void**)(&api.n) = dlsym(api.lib, "nccl" #n)
  RFX_SYM(GetUniqueId); RFX_SYM(CommInitRank); RFX_SYM(CommDestroy); RFX_SYM(AllGather); RFX_SYM(Broadcast); RFX_SYM(GroupStart); RFX_SYM(GroupEnd);
  RFX_SYM(GetErrorString);
#undef RFX_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.GroupStart && api.GroupEnd && api.GetErrorString;
  return &api;
}

__global__ void stamp_kernel(unsigned long long* t) {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  *t = v;
}
__global__ void elapsed_kernel(const unsigned long long* t0, float* ms) {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  *ms = (float)((double)(v - *t0) * 1e-6);
}

// Frame barrier through peer memory (RFX_GROUP_BARRIER=flags): after its last kernel every rank stores (frame, kernel ms) into its
// slot of EVERY rank's flag array (peer stores over NVLink); the next thing on each rank's stream spins until all slots carry the
// frame.  Same information as the NCCL all-gather, without a collective launch.  A rank that never arrives makes the waiters give
// up after ~2^32 cycles and raise `err` instead of hanging the GPU.
struct SyncPeers { unsigned long long* slots[RFX_MAX_PEERS]; int n; };
__global__ void publish_kernel(const unsigned long long* t0, float* ms_out, SyncPeers sp, int rank, unsigned frame) {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  const float ms = (float)((double)(v - *t0) * 1e-6);
  *ms_out = ms;
  const unsigned long long word = ((unsigned long long)frame << 32) | (unsigned long long)__float_as_uint(ms);
  __threadfence_system();  // this rank's frame (earlier kernels on the stream) is visible before the flag is
  // two slot sets by frame parity: a rank can publish frame f + 1 before a slow peer has READ its frame-f word, but not f + 2
  // (that needs the peer's own f + 1 flag, which the peer sets after its frame-f wait) - so every rank reads the same costs
  const int set = (int)(frame & 1u) * RFX_MAX_PEERS;
  for (int p = 0; p < sp.n; p++) *((volatile unsigned long long*)(sp.slots[p] + set + rank)) = word;
}
__global__ void wait_flags_kernel(volatile unsigned long long* slots, int n, unsigned frame, float* ms_all, int* err) {
  const int r = threadIdx.x;
  if (r < n) {
    slots += (frame & 1u) * RFX_MAX_PEERS;
    const long long start = clock64();
    unsigned long long v = slots[r];
    while ((unsigned)(v >> 32) != frame) {
      if (clock64() - start > (1LL << 32)) { *err = 1; break; }
      v = slots[r];
    }
    ms_all[r] = __uint_as_float((unsigned)(v & 0xffffffffull));
  }
  __threadfence_system();
}

}  // namespace

// ---- sharding plan (pure host arithmetic; mirrored by realism_effects_b200/parallel.py:ShardPlan for the CPU tests) -------------
// Output rows [a,b) of every launch of one frame (chain order: K1, K2, K3 pass 0..n-1, K4) for the band [own0, own1):
// launch k runs on the band widened by the rows all later launches read around their outputs.
extern "C" rfx_status rfx_shard_ranges(uint32_t width, uint32_t height, uint32_t own0, uint32_t own1, int32_t n_poisson_passes, float radius, int32_t ssgi_mode,
                                       uint32_t* ranges, uint32_t n_launches) {
  (void)ssgi_mode;  // K4 runs in both modes (SSR composes with inputType specular, DenoiserComposePass.js:26-33)
  const uint32_t expect = 3u + (uint32_t)n_poisson_passes;
  if (!ranges || n_launches != expect || own0 >= own1 || own1 > height || width == 0 || n_poisson_passes < 0) return RFX_ERR_INVALID_ARG;
  const int H = (int)height;
  // a Poisson tap offset is rotated AFTER the division by the resolution (poisson_denoise.frag:183-189), so its row reach is
  // radius * max(1, H/W); + 1 row for the bilinear footprint / the quad-derivative helper row
  const int poisson_halo = (int)std::ceil((double)radius * std::max(1.0, (double)height / (double)width)) + 1;
  const int k2_rows = 2;  // 5x5 clamp window (reproject.frag:57-59)
  const int k4_rows = 1;  // the exact K4's literal bilinear fetch at the pixel centre
  auto expand = [&](int& a, int& b, int rows) { a = std::max(0, a - rows); b = std::min(H, b + rows); };
  const int n = n_poisson_passes;
  std::vector<int> r0(expect), r1(expect);
  int a = (int)own0, b = (int)own1;
  r0[expect - 1] = a; r1[expect - 1] = b;
  if (n) {
    expand(a, b, k4_rows);
    r0[2 + n - 1] = a; r1[2 + n - 1] = b;
    for (int j = n - 2; j >= 0; j--) { expand(a, b, poisson_halo); r0[2 + j] = a; r1[2 + j] = b; }
    expand(a, b, poisson_halo);
  }
  r0[1] = a; r1[1] = b;
  expand(a, b, k2_rows);
  r0[0] = a; r1[0] = b;
  for (uint32_t k = 0; k < expect; k++) { ranges[2 * k] = (uint32_t)r0[k]; ranges[2 * k + 1] = (uint32_t)r1[k]; }
  return RFX_OK;
}

// New band borders from per-rank costs measured with `measured` borders: cost per row constant inside each measured band, ideal
// border k where the cumulative cost reaches k/N of the total; move `damping` of the way there, align to 16 rows, keep every band
// between min_rows and max_share x the mean height (ALL bands, the last one included).  Deterministic: every rank computes the
// same borders from the same gathered costs.
extern "C" rfx_status rfx_shard_rebalance(const uint32_t* bounds, const uint32_t* measured, const float* costs, int32_t n, uint32_t* out) {
  if (!bounds || !costs || !out || n < 1) return RFX_ERR_INVALID_ARG;
  const uint32_t* mb = measured ? measured : bounds;
  const int H = (int)bounds[n], align = 16, min_rows = 64;
  const double max_share = 4.0, damping = 0.6;
  std::vector<double> c(n);
  double total = 0.0;
  for (int i = 0; i < n; i++) { c[i] = std::max((double)costs[i], 1e-9); total += c[i]; }
  std::vector<double> ideal(n + 1, 0.0);
  int k = 0;
  double acc = 0.0;
  for (int i = 1; i < n; i++) {
    const double want = total * i / n;
    while (k < n - 1 && acc + c[k] < want) { acc += c[k]; k++; }
    ideal[i] = mb[k] + (want - acc) / c[k] * ((double)mb[k + 1] - (double)mb[k]);
  }
  ideal[n] = H;
  const int lo_h = min_rows, hi_h = std::max(min_rows, (int)(max_share * H / n));
  std::vector<int> o(n + 1, 0);
  for (int i = 1; i < n; i++) {
    double b = bounds[i] + damping * (ideal[i] - (double)bounds[i]);
    int bi = (int)std::lround(b / align) * align;
    bi = std::max(bi, o[i - 1] + lo_h);
    bi = std::min(bi, o[i - 1] + hi_h);
    bi = std::min(bi, H - (n - i) * lo_h);
    o[i] = bi;
  }
  o[n] = H;
  for (int i = n - 1; i >= 1; i--)  // the cap holds for the bands below too: push borders down so that no band exceeds hi_h
    if (o[i + 1] - o[i] > hi_h) o[i] = std::min(o[i + 1] - lo_h, ((o[i + 1] - hi_h + align - 1) / align) * align);
  for (int i = 0; i <= n; i++) out[i] = (uint32_t)o[i];
  return RFX_OK;
}

#define RFX_GROUP_RING 8
struct rfx_group {
  rfx_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  rfx_ssgi_chain* chain = nullptr;
  std::vector<void*> opened;                 // peer mappings to close
  std::vector<uint32_t> bounds;              // bands of the frame being rendered (world + 1)
  std::vector<uint32_t> bounds_ring[RFX_GROUP_RING];  // bands each recent frame was rendered with
  uint64_t frame = 0;
  bool began = false;  // begin_frame already ran for the frame about to be rendered
  bool use_peer = true;  // history rows read in place on their owner (CUDA IPC); false: replicated by an NCCL exchange after every frame
  bool inprocess = false;  // every member lives in this process: plain device pointers, no NCCL, the host orders the frames
  bool use_flags = false;  // frame barrier through peer-memory flags instead of the NCCL all-gather (RFX_GROUP_BARRIER=flags; needs use_peer)
  unsigned long long* d_sync = nullptr;  // [2][RFX_MAX_PEERS] slots by frame parity: rank r's (frame, ms) word, written by rank r
  SyncPeers sync_peers{};
  int* d_err = nullptr;
  // device-timed cost of this rank's kernels, all-gathered every frame (the collective doubles as the frame barrier)
  unsigned long long* d_t0 = nullptr;
  float* d_ms = nullptr;     // [1 + world]: own, then everyone's
  float* h_ms = nullptr;     // pinned ring [RFX_GROUP_RING][world]
  cudaEvent_t ev[RFX_GROUP_RING]{};
  bool ev_valid[RFX_GROUP_RING]{};
  int rebalance_every = 0, rebalance_lag = 2;
  float last_costs[RFX_MAX_PEERS]{};
};

#define NC(call)                                                                                                      \
  do {                                                                                                                \
    ncclResult_t r_ = (call);                                                                                         \
    if (r_ != ncclSuccess) return fail(ctx, RFX_ERR_NCCL, "%s failed: %s", #call, nccl_api()->GetErrorString(r_));    \
  } while (0)

extern "C" {

rfx_status rfx_group_get_unique_id(void* id128) {
  if (!id128) return RFX_ERR_INVALID_ARG;
  NcclApi* n = nccl_api();
  if (!n->ok) return RFX_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == RFX_GROUP_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return RFX_ERR_NCCL;
  memcpy(id128, &id, sizeof id);
  return RFX_OK;
}

rfx_status rfx_group_create(rfx_ctx* ctx, const void* id128, int32_t rank, int32_t world, rfx_group** out) {
  if (!ctx || !id128 || !out || world < 1 || world > RFX_MAX_PEERS || rank < 0 || rank >= world) return fail(ctx, RFX_ERR_INVALID_ARG, "group_create: bad arguments (world <= %d)", RFX_MAX_PEERS);
  *out = nullptr;
  NcclApi* n = nccl_api();
  if (!n->ok) return fail(ctx, RFX_ERR_NCCL, "group_create: libnccl.so.2 could not be loaded");
  CU(cudaSetDevice(ctx->device));
  rfx_group* g = new rfx_group();
  g->ctx = ctx; g->rank = rank; g->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclResult_t r = n->CommInitRank(&g->comm, world, id, rank);
  if (r != ncclSuccess) { delete g; return fail(ctx, RFX_ERR_NCCL, "ncclCommInitRank failed: %s", n->GetErrorString(r)); }
  bool ok = cudaMalloc(&g->d_t0, 8) == cudaSuccess && cudaMalloc(&g->d_ms, sizeof(float) * (1 + world)) == cudaSuccess &&
            cudaHostAlloc(&g->h_ms, sizeof(float) * RFX_GROUP_RING * world, cudaHostAllocDefault) == cudaSuccess;
  for (int i = 0; i < RFX_GROUP_RING && ok; i++) ok = cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { rfx_group_destroy(g); return fail(ctx, RFX_ERR_CUDA, "group_create: allocation failed"); }
  *out = g;
  return RFX_OK;
}

rfx_status rfx_group_create_inprocess(rfx_ctx* ctx, int32_t rank, int32_t world, rfx_group** out) {
  if (!ctx || !out || world < 1 || world > RFX_MAX_PEERS || rank < 0 || rank >= world) return fail(ctx, RFX_ERR_INVALID_ARG, "group_create_inprocess: bad arguments (world <= %d)", RFX_MAX_PEERS);
  *out = nullptr;
  CU(cudaSetDevice(ctx->device));
  rfx_group* g = new rfx_group();
  g->ctx = ctx; g->rank = rank; g->world = world; g->inprocess = true;
  bool ok = cudaMalloc(&g->d_t0, 8) == cudaSuccess && cudaMalloc(&g->d_ms, sizeof(float) * (1 + world)) == cudaSuccess &&
            cudaMemset(g->d_ms, 0, sizeof(float) * (1 + world)) == cudaSuccess &&
            cudaHostAlloc(&g->h_ms, sizeof(float) * RFX_GROUP_RING * world, cudaHostAllocDefault) == cudaSuccess;
  for (int i = 0; i < RFX_GROUP_RING && ok; i++) ok = cudaEventCreateWithFlags(&g->ev[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { rfx_group_destroy(g); return fail(ctx, RFX_ERR_CUDA, "group_create_inprocess: allocation failed"); }
  *out = g;
  return RFX_OK;
}

// Attaches chains[r] to groups[r] for every member at once: the peer tables hold the members' own device pointers.
rfx_status rfx_group_attach_chains_inprocess(rfx_group* const* groups, rfx_ssgi_chain* const* chains, int32_t world) {
  if (!groups || !chains || world < 1 || world > RFX_MAX_PEERS) return RFX_ERR_INVALID_ARG;
  for (int r = 0; r < world; r++) {
    rfx_group* g = groups[r];
    rfx_ssgi_chain* ch = chains[r];
    if (!g || !ch) return RFX_ERR_INVALID_ARG;
    rfx_ctx* ctx = g->ctx;
    if (!g->inprocess || g->world != world || g->rank != r) return fail(ctx, RFX_ERR_INVALID_ARG, "group_attach_chains_inprocess: member %d is not an in-process member of rank %d / world %d", r, r, world);
    if (ch->ctx != ctx) return fail(ctx, RFX_ERR_INVALID_ARG, "group_attach_chains_inprocess: chain %d belongs to another context", r);
    if (!ch->fastpath) return fail(ctx, RFX_ERR_UNSUPPORTED, "group_attach_chains_inprocess: row-sharded groups need the fast SSGI chain (fast_math on, mode SSGI)");
    if (g->chain || ch->group) return fail(ctx, RFX_ERR_INVALID_ARG, "group_attach_chains_inprocess: already attached");
    if (ch->opt.width != chains[0]->opt.width || ch->opt.height != chains[0]->opt.height) return fail(ctx, RFX_ERR_SIZE_MISMATCH, "group_attach_chains_inprocess: chain sizes differ");
    if ((int)ch->opt.height < world * 64) return fail(ctx, RFX_ERR_UNSUPPORTED, "group_attach_chains_inprocess: need at least 64 rows per rank");
  }
  const int W = (int)chains[0]->opt.width, H = (int)chains[0]->opt.height, n = world;
  for (int r = 0; r < n; r++)  // members on different devices dereference each other's pointers: peer access both ways
    for (int q = 0; q < n; q++) {
      const int dr = groups[r]->ctx->device, dq = groups[q]->ctx->device;
      if (dr == dq) continue;
      rfx_ctx* ctx = groups[r]->ctx;
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, dr, dq));
      if (!can) return fail(ctx, RFX_ERR_UNSUPPORTED, "group_attach_chains_inprocess: device %d cannot access device %d (use one process per GPU: rfx_group_create)", dr, dq);
      CU(cudaSetDevice(dr));
      const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, RFX_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dr, dq, cudaGetErrorString(e));
      cudaGetLastError();
    }
  for (int r = 0; r < n; r++) {
    rfx_group* g = groups[r];
    rfx_ssgi_chain* ch = chains[r];
    for (int b = 0; b < 2; b++) {
      PeerPV& pc = ch->peer_composed[b];
      pc = PeerPV{};
      pc.local = PV{(const unsigned char*)ch->composed2[b].ptr, W, H, (long long)ch->composed2[b].pitch};
      PeerPV& pd = ch->peer_dn[b];
      pd = PeerPV{};
      pd.local = PV{(const unsigned char*)ch->dnB16[b].p, W, H, (long long)ch->dnB16[b].pitch};
      PeerPV& pa = ch->peer_dnA[b];
      pa = PeerPV{};
      pa.local = PV{(const unsigned char*)ch->dnA16[b].p, W, H, (long long)ch->dnA16[b].pitch};
      pc.n = pd.n = pa.n = n;
      for (int q = 0; q < n; q++) {
        pc.base[q] = (const unsigned char*)chains[q]->composed2[b].ptr; pd.base[q] = (const unsigned char*)chains[q]->dnB16[b].p;
        pa.base[q] = (const unsigned char*)chains[q]->dnA16[b].p;
      }
    }
    g->use_peer = true;
    g->use_flags = false;
    g->bounds.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; i++) g->bounds[i] = (uint32_t)(std::lround((double)H * i / n / 16.0) * 16);
    g->bounds[n] = (uint32_t)H;
    for (auto& bb : g->bounds_ring) bb = g->bounds;
    g->chain = ch;
    ch->group = g;
  }
  return RFX_OK;
}

void rfx_group_destroy(rfx_group* g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaDeviceSynchronize();
  if (g->chain) { g->chain->group = nullptr; g->chain = nullptr; }
  for (void* p : g->opened) cudaIpcCloseMemHandle(p);
  if (g->comm) nccl_api()->CommDestroy(g->comm);
  cudaFree(g->d_t0); cudaFree(g->d_ms); cudaFree(g->d_sync); cudaFree(g->d_err);
  if (g->h_ms) cudaFreeHost(g->h_ms);
  for (cudaEvent_t e : g->ev) if (e) cudaEventDestroy(e);
  delete g;
}

int32_t rfx_group_rank(const rfx_group* g) { return g ? g->rank : -1; }
int32_t rfx_group_world(const rfx_group* g) { return g ? g->world : 0; }
int32_t rfx_group_uses_peer_reads(const rfx_group* g) { return g && g->use_peer ? 1 : 0; }

// Collective.  Attaches a fast SSGI chain (same options on every rank) to the group: every rank exports its double-buffered
// `composed` and `dn` planes (CUDA IPC), the handles are all-gathered and every peer's planes are mapped here.  Bands start equal.
rfx_status rfx_group_attach_chain(rfx_group* g, rfx_ssgi_chain* ch) {
  if (!g || !ch) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = g->ctx;
  if (ch->ctx != ctx) return fail(ctx, RFX_ERR_INVALID_ARG, "group_attach_chain: the chain belongs to another context");
  if (!ch->fastpath) return fail(ctx, RFX_ERR_UNSUPPORTED, "group_attach_chain: row-sharded groups need the fast SSGI chain (fast_math on, mode SSGI)");
  if (g->chain || ch->group) return fail(ctx, RFX_ERR_INVALID_ARG, "group_attach_chain: already attached");
  const int W = (int)ch->opt.width, H = (int)ch->opt.height, n = g->world;
  if (H < n * 64) return fail(ctx, RFX_ERR_UNSUPPORTED, "group_attach_chain: need at least 64 rows per rank");
  CU(cudaSetDevice(ctx->device));
  if (!g->d_sync) {
    CU(cudaMalloc(&g->d_sync, sizeof(unsigned long long) * 2 * RFX_MAX_PEERS));
    CU(cudaMemset(g->d_sync, 0, sizeof(unsigned long long) * 2 * RFX_MAX_PEERS));
    CU(cudaMalloc(&g->d_err, sizeof(int)));
    CU(cudaMemset(g->d_err, 0, sizeof(int)));
  }
  constexpr int NH = 7;  // exported allocations per rank: composed x2, dn (B) x2, the flag array, dn (A) x2
  void* mine[NH] = {ch->composed2[0].ptr, ch->composed2[1].ptr, ch->dnB16[0].p, ch->dnB16[1].p, g->d_sync, ch->dnA16[0].p, ch->dnA16[1].p};
  std::vector<void*> all((size_t)NH * n, nullptr);
  if (n > 1) {
    std::vector<cudaIpcMemHandle_t> hs((size_t)NH * n);
    cudaIpcMemHandle_t my[NH];
    for (int i = 0; i < NH; i++) CU(cudaIpcGetMemHandle(&my[i], mine[i]));
    unsigned char* d = nullptr;
    CU(cudaMalloc(&d, sizeof(my) * (size_t)(n + 1)));
    CU(cudaMemcpy(d, my, sizeof my, cudaMemcpyHostToDevice));
    NC(nccl_api()->AllGather(d, d + sizeof my, sizeof my, ncclChar, g->comm, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(hs.data(), d + sizeof my, sizeof(my) * (size_t)n, cudaMemcpyDeviceToHost));
    CU(cudaFree(d));
    int ok = 1;
    if (const char* e = getenv("RFX_GROUP_EXCHANGE")) ok = strcmp(e, "allgather") != 0;  // force the replicated fallback (A/B measurements)
    for (int r = 0; r < n && ok; r++)
      for (int i = 0; i < NH && ok; i++) {
        if (r == g->rank) { all[(size_t)r * NH + i] = mine[i]; continue; }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hs[(size_t)r * NH + i], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { ok = 0; cudaGetLastError(); ctx->err = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e); break; }
        g->opened.push_back(p);
        all[(size_t)r * NH + i] = p;
      }
    // every rank must take the same path: all-gather the flags (device bounce) and AND them
    int* dflag = nullptr;
    CU(cudaMalloc(&dflag, sizeof(int) * (size_t)(n + 1)));
    CU(cudaMemcpy(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NC(nccl_api()->AllGather(dflag, dflag + 1, sizeof(int), ncclChar, g->comm, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<int> flags((size_t)n);
    CU(cudaMemcpy(flags.data(), dflag + 1, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    CU(cudaFree(dflag));
    for (int r = 0; r < n; r++) ok = ok && flags[(size_t)r];
    g->use_peer = ok != 0;
    if (!g->use_peer) {  // replicated fallback: every rank keeps full copies, exchanged after every frame (rfx_group_allgather_rows)
      for (void* p : g->opened) cudaIpcCloseMemHandle(p);
      g->opened.clear();
      for (int r = 0; r < n; r++) for (int i = 0; i < NH; i++) all[(size_t)r * NH + i] = mine[i];
    }
  } else {
    for (int i = 0; i < NH; i++) all[i] = mine[i];
  }
  for (int b = 0; b < 2; b++) {
    PeerPV& pc = ch->peer_composed[b];
    pc = PeerPV{};
    pc.local = PV{(const unsigned char*)ch->composed2[b].ptr, W, H, (long long)ch->composed2[b].pitch};
    PeerPV& pd = ch->peer_dn[b];
    pd = PeerPV{};
    pd.local = PV{(const unsigned char*)ch->dnB16[b].p, W, H, (long long)ch->dnB16[b].pitch};
    PeerPV& pa = ch->peer_dnA[b];
    pa = PeerPV{};
    pa.local = PV{(const unsigned char*)ch->dnA16[b].p, W, H, (long long)ch->dnA16[b].pitch};
    pc.n = pd.n = pa.n = g->use_peer ? n : 1;
    for (int r = 0; r < n; r++) {
      pc.base[r] = (const unsigned char*)all[(size_t)r * NH + b]; pd.base[r] = (const unsigned char*)all[(size_t)r * NH + 2 + b];
      pa.base[r] = (const unsigned char*)all[(size_t)r * NH + 5 + b];
    }
  }
  g->sync_peers.n = n;
  for (int r = 0; r < n; r++) g->sync_peers.slots[r] = (unsigned long long*)all[(size_t)r * NH + 4];
  g->use_flags = false;
  if (const char* e = getenv("RFX_GROUP_BARRIER")) g->use_flags = g->use_peer && n > 1 && strcmp(e, "flags") == 0;
  g->bounds.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; i++) g->bounds[i] = (uint32_t)(std::lround((double)H * i / n / 16.0) * 16);
  g->bounds[n] = (uint32_t)H;
  for (auto& b : g->bounds_ring) b = g->bounds;
  g->chain = ch;
  ch->group = g;
  return RFX_OK;
}

rfx_status rfx_group_get_bounds(const rfx_group* g, uint32_t* bounds) {
  if (!g || !bounds || g->bounds.empty()) return RFX_ERR_INVALID_ARG;
  memcpy(bounds, g->bounds.data(), sizeof(uint32_t) * g->bounds.size());
  return RFX_OK;
}
// Every rank must pass the same ascending borders (bounds[0] = 0, bounds[world] = height); they apply from the next frame on.
rfx_status rfx_group_set_bounds(rfx_group* g, const uint32_t* bounds) {
  if (!g || !bounds || !g->chain) return RFX_ERR_INVALID_ARG;
  const int n = g->world;
  if (bounds[0] != 0 || bounds[n] != g->chain->opt.height) return fail(g->ctx, RFX_ERR_INVALID_ARG, "group_set_bounds: borders must run from 0 to the frame height");
  for (int i = 0; i < n; i++) if (bounds[i] >= bounds[i + 1]) return fail(g->ctx, RFX_ERR_INVALID_ARG, "group_set_bounds: borders must ascend");
  g->bounds.assign(bounds, bounds + n + 1);
  return RFX_OK;
}
// every > 0: every `every` frames the borders move towards equal device-timed cost (all ranks decide from the same gathered
// times, `lag` frames old so that nobody waits for them); 0: static bands.
rfx_status rfx_group_set_rebalance(rfx_group* g, int32_t every, int32_t lag) {
  if (!g || every < 0 || lag < 1 || lag >= RFX_GROUP_RING) return RFX_ERR_INVALID_ARG;
  g->rebalance_every = every; g->rebalance_lag = lag;
  return RFX_OK;
}
rfx_status rfx_group_last_costs(const rfx_group* g, float* ms) {
  if (!g || !ms) return RFX_ERR_INVALID_ARG;
  memcpy(ms, g->last_costs, sizeof(float) * (size_t)g->world);
  return RFX_OK;
}

// In lockstep on every rank (no communication): applies the cost-driven border move that is due and returns the borders the next
// frame will be rendered with.  rfx_ssgi_chain_render_sharded calls it implicitly; a host path calls it first to size its uploads.
rfx_status rfx_group_begin_frame(rfx_group* g, uint32_t* bounds_out) {
  if (!g || !g->chain) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = g->ctx;
  const int n = g->world;
  if (!g->began) {
    g->began = true;
    // cost-driven borders (deterministic on every rank: same gathered times, same arithmetic)
    if (!g->inprocess && g->rebalance_every > 0 && g->frame >= (uint64_t)g->rebalance_lag && g->frame % (uint64_t)g->rebalance_every == 0) {
      const uint64_t src = g->frame - (uint64_t)g->rebalance_lag;
      const int slot = (int)(src % RFX_GROUP_RING);
      if (g->ev_valid[slot]) {
        CU(cudaEventSynchronize(g->ev[slot]));
        std::vector<uint32_t> nb((size_t)n + 1);
        memcpy(g->last_costs, g->h_ms + (size_t)slot * n, sizeof(float) * (size_t)n);
        if (rfx_shard_rebalance(g->bounds.data(), g->bounds_ring[slot].data(), g->last_costs, n, nb.data()) == RFX_OK) g->bounds = nb;
      }
    }
  }
  if (bounds_out) memcpy(bounds_out, g->bounds.data(), sizeof(uint32_t) * g->bounds.size());
  return RFX_OK;
}
rfx_status rfx_group_get_last_bounds(const rfx_group* g, uint32_t* bounds) {  // the borders the most recent frame was rendered with
  if (!g || !bounds || g->bounds.empty()) return RFX_ERR_INVALID_ARG;
  const std::vector<uint32_t>& b = g->frame ? g->bounds_ring[(g->frame - 1) % RFX_GROUP_RING] : g->bounds;
  memcpy(bounds, b.data(), sizeof(uint32_t) * b.size());
  return RFX_OK;
}

// Collective: completes a full-frame INPUT plane of which every rank has uploaded only its own rows [bounds[r], bounds[r+1]):
// one NCCL group of per-rank broadcasts (an all-gather with unequal counts), in place, over NVLink.  This is the path's one
// real exchange of the host-buffer route: depth and velocity are sampled at arbitrary screen positions by every rank.
rfx_status rfx_group_allgather_rows(rfx_group* g, void* stream, const rfx_plane* plane, const uint32_t* bounds) {
  if (!g || !plane || !plane->ptr || !bounds) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = g->ctx;
  const cudaStream_t cs = stream ? (cudaStream_t)stream : ctx->stream;
  if (bounds[g->world] > plane->height) return fail(ctx, RFX_ERR_INVALID_ARG, "group_allgather_rows: borders exceed the plane");
  if (g->world == 1) return RFX_OK;
  NcclApi* n = nccl_api();
  NC(n->GroupStart());
  for (int r = 0; r < g->world; r++) {
    unsigned char* p = (unsigned char*)plane->ptr + (size_t)bounds[r] * plane->pitch;
    const size_t bytes = (size_t)(bounds[r + 1] - bounds[r]) * plane->pitch;
    if (bytes) NC(n->Broadcast(p, p, bytes, ncclChar, r, g->comm, cs));
  }
  NC(n->GroupEnd());
  return RFX_OK;
}

// Collective: one frame, this rank's band.  Inputs are full-frame planes (every rank holds them).  Ends with the group's
// per-frame collective on `stream`, after which every rank's rows of this frame are visible to its peers.
rfx_status rfx_ssgi_chain_render_sharded(rfx_ssgi_chain* ch, void* stream, const rfx_ssgi_frame* f) {
  if (!ch || !f) return RFX_ERR_INVALID_ARG;
  rfx_ctx* ctx = ch->ctx;
  rfx_group* g = ch->group;
  if (!g) return fail(ctx, RFX_ERR_NOT_READY, "render_sharded: the chain is not attached to a group (rfx_group_attach_chain)");
  const int n = g->world;
  const cudaStream_t cs = stream ? (cudaStream_t)stream : ctx->stream;
  rfx_status st = rfx_group_begin_frame(g, nullptr);
  if (st != RFX_OK) return st;
  g->began = false;
  const int cur = (int)(ch->frame_idx & 1), prev = cur ^ 1;
  // owners of the data this frame READS (last frame's bands) and of the rows it carries forward
  const std::vector<uint32_t>& pb = g->bounds_ring[(g->frame + RFX_GROUP_RING - 1) % RFX_GROUP_RING];
  for (PeerPV* p : {&ch->peer_composed[prev], &ch->peer_dn[prev], &ch->peer_dnA[prev]}) {
    for (int i = 0; i <= n; i++) p->bound[i] = (int)pb[i];
    p->own0 = (int)pb[g->rank]; p->own1 = (int)pb[g->rank + 1];
    if (g->frame == 0 || !g->use_peer) { p->own0 = 0; p->own1 = (int)ch->opt.height; }  // nothing rendered yet (all zero) / replicated planes: every row is local
  }
  const uint32_t n_launches = 3u + 2u * (uint32_t)ch->opt.denoise_iterations;
  std::vector<uint32_t> ranges((size_t)n_launches * 2);
  st = rfx_shard_ranges(ch->opt.width, ch->opt.height, g->bounds[g->rank], g->bounds[g->rank + 1], 2 * ch->opt.denoise_iterations, ch->opt.radius, 1, ranges.data(), n_launches);
  if (st != RFX_OK) return fail(ctx, st, "render_sharded: bad band");
  stamp_kernel<<<1, 1, 0, cs>>>(g->d_t0);
  if ((st = chain_render_impl(ch, stream, f, ranges.data(), 1, 0, 0xffffffffu)) != RFX_OK) return st;
  if (g->use_flags) {  // barrier + cost exchange through peer-memory flags
    publish_kernel<<<1, 1, 0, cs>>>(g->d_t0, g->d_ms, g->sync_peers, g->rank, (unsigned)(g->frame + 1));
    wait_flags_kernel<<<1, 32, 0, cs>>>(g->d_sync, n, (unsigned)(g->frame + 1), g->d_ms + 1, g->d_err);
    ctx->launches += 3;
  } else {
    elapsed_kernel<<<1, 1, 0, cs>>>(g->d_t0, g->d_ms);
    ctx->launches += 2;
  }
  if (!g->use_peer && n > 1) {  // replicated fallback: every rank's rows of this frame's history planes to every other rank (32 B/px of the whole frame)
    rfx_plane dnp{};
    dnp.ptr = ch->dnB16[cur].p; dnp.width = ch->opt.width; dnp.height = ch->opt.height; dnp.pitch = ch->dnB16[cur].pitch; dnp.format = RFX_FMT_RGBA32F;
    if ((st = rfx_group_allgather_rows(g, stream, &ch->composed2[cur], g->bounds.data())) != RFX_OK) return st;
    if ((st = rfx_group_allgather_rows(g, stream, &dnp, g->bounds.data())) != RFX_OK) return st;
    dnp.ptr = ch->dnA16[cur].p; dnp.pitch = ch->dnA16[cur].pitch;  // the A target's rows too: discarded texels are carried from them (chain_render_fast)
    if ((st = rfx_group_allgather_rows(g, stream, &dnp, g->bounds.data())) != RFX_OK) return st;
  }
  const int slot = (int)(g->frame % RFX_GROUP_RING);
  if (g->inprocess) CU(cudaMemcpyAsync(g->d_ms + 1 + g->rank, g->d_ms, sizeof(float), cudaMemcpyDeviceToDevice, cs));  // the host orders the members' frames; only this member's cost is known here
  else if (!g->use_flags) NC(nccl_api()->AllGather(g->d_ms, g->d_ms + 1, 1, ncclFloat, g->comm, cs));  // every rank's frame is complete when this completes
  CU(cudaMemcpyAsync(g->h_ms + (size_t)slot * n, g->d_ms + 1, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, cs));
  CU(cudaEventRecord(g->ev[slot], cs));
  g->ev_valid[slot] = true;
  g->bounds_ring[slot] = g->bounds;
  g->frame++;
  return RFX_OK;
}

}  // extern "C"
