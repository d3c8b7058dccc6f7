// C ABI of the code encoder (include/lvs.h, "embedding on the GPUs that search": SURVEY section 8f row 4).
//
// The reference embeds a chunk with UniXcoder - transformers' RobertaModel over the token ids, masked mean pooling (reference
// src/lattice/providers/unixcoder_provider.py:137-155) - on whatever device torch finds, converts the result to python lists
// (:194-215) and hands them to QdrantManager.upsert (embeddings/indexer.py:77-86).  Here the forward pass runs as hand-written
// kernels (linear_kernel.cuh: tcgen05 GEMMs with fused bias / GELU / residual epilogues; encoder_kernels.cuh: embeddings +
// LayerNorm, attention, LayerNorm, pooling) and lvs_encoder_embed_upsert feeds the pooled vectors straight into the shard's upsert
// kernel: they never leave HBM.
#include "../../include/lvs.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "encoder_kernels.cuh"
#include "host_util.h"
#include "linear_kernel.cuh"

using namespace lvs;

namespace {

struct Dense {
    __nv_bfloat16* w = nullptr;   // [out][in] bf16
    float* b = nullptr;           // [out]
    int out = 0, in = 0;
};
struct Layer {
    Dense qkv, attn_out, inter, out;
    float* ln1_w = nullptr; float* ln1_b = nullptr; float* ln2_w = nullptr; float* ln2_b = nullptr;
};

__global__ void f32_to_bf16_kernel(const float* src, __nv_bfloat16* dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16(src[i]);
}

}  // namespace

struct lvs_encoder {
    lvs_encoder_config cfg;
    float* word = nullptr; float* pos = nullptr; float* type0 = nullptr; float* eln_w = nullptr; float* eln_b = nullptr;
    std::vector<Layer> layers;
    std::vector<std::string> loaded;
    size_t n_expected = 0;
    // workspace (grown on demand)
    int64_t ws_tokens = 0;
    __nv_bfloat16* x = nullptr; __nv_bfloat16* qkv = nullptr; __nv_bfloat16* ctx = nullptr; __nv_bfloat16* h = nullptr;
    float* y = nullptr; float* pooled = nullptr; int32_t* d_ids = nullptr; int32_t* d_err = nullptr;
    int64_t ws_seqs = 0;
    void* stage = nullptr; size_t stage_bytes = 0;     // pinned staging for weights / ids
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0.f;
    int opt_pair = 1;          // 0: dense layers never use the CTA-pair form (env LATTICE_B200_ENCODER_PAIR=0; for A/B timing)
    std::mutex mu;
};

static int dev_alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return lvs_fail(LVS_ENOMEM, "encoder: cannot allocate %zu bytes on the device: %s", bytes, cudaGetErrorString(e)); }
    return LVS_OK;
}

extern "C" int lvs_encoder_create(const lvs_encoder_config* cfg, lvs_encoder** out) {
    lvs_lib_bind_thread();
    if (!lvs_lib_ready()) return lvs_fail(LVS_ESTATE, "lvs_init() has not been called (no CUDA device bound)");
    if (!cfg || !out) return lvs_fail(LVS_EINVAL, "NULL argument");
    *out = nullptr;
    if (cfg->hidden < 64 || cfg->hidden > 1024 || cfg->hidden % 64 != 0) return lvs_fail(LVS_ELIMIT, "hidden size %d must be a multiple of 64 in 64..1024", cfg->hidden);
    if (cfg->n_heads < 1 || cfg->hidden != cfg->n_heads * kAttnD) return lvs_fail(LVS_ELIMIT, "attention heads must be %d wide (hidden %d, heads %d)", kAttnD, cfg->hidden, cfg->n_heads);
    if (cfg->intermediate < 64 || cfg->intermediate % 64 != 0) return lvs_fail(LVS_ELIMIT, "intermediate size %d must be a multiple of 64", cfg->intermediate);
    if (cfg->n_layers < 1 || cfg->n_layers > 64 || cfg->vocab < 2 || cfg->max_pos < 4) return lvs_fail(LVS_EINVAL, "bad layer count / vocabulary / positions");
    lvs_encoder* e = new (std::nothrow) lvs_encoder();
    if (!e) return lvs_fail(LVS_ENOMEM, "host allocation failed");
    e->cfg = *cfg;
    if (const char* ev = getenv("LATTICE_B200_ENCODER_PAIR")) e->opt_pair = atoi(ev) != 0;
    const int H = cfg->hidden, I = cfg->intermediate;
    int rc = LVS_OK;
    auto A = [&](void** p, size_t bytes) { if (rc == LVS_OK) rc = dev_alloc(p, bytes); };
    A((void**)&e->word, (size_t)cfg->vocab * H * 4); A((void**)&e->pos, (size_t)cfg->max_pos * H * 4); A((void**)&e->type0, (size_t)H * 4);
    A((void**)&e->eln_w, (size_t)H * 4); A((void**)&e->eln_b, (size_t)H * 4); A((void**)&e->d_err, 4);
    e->layers.resize(cfg->n_layers);
    for (auto& L : e->layers) {
        L.qkv.out = 3 * H; L.qkv.in = H; L.attn_out.out = H; L.attn_out.in = H; L.inter.out = I; L.inter.in = H; L.out.out = H; L.out.in = I;
        for (Dense* d : {&L.qkv, &L.attn_out, &L.inter, &L.out}) { A((void**)&d->w, (size_t)d->out * d->in * 2); A((void**)&d->b, (size_t)d->out * 4); }
        A((void**)&L.ln1_w, (size_t)H * 4); A((void**)&L.ln1_b, (size_t)H * 4); A((void**)&L.ln2_w, (size_t)H * 4); A((void**)&L.ln2_b, (size_t)H * 4);
    }
    e->n_expected = 5 + (size_t)cfg->n_layers * 16;
    if (rc == LVS_OK && cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) rc = lvs_fail(LVS_ECUDA, "encoder: stream creation failed");
    if (rc == LVS_OK && (cudaEventCreate(&e->ev[0]) != cudaSuccess || cudaEventCreate(&e->ev[1]) != cudaSuccess)) rc = lvs_fail(LVS_ECUDA, "encoder: event creation failed");
    if (rc == LVS_OK && cudaMemset(e->d_err, 0, 4) != cudaSuccess) rc = lvs_fail(LVS_ECUDA, "encoder: memset failed");
    if (rc != LVS_OK) { lvs_encoder_destroy(e); return rc; }
    *out = e;
    return LVS_OK;
}

extern "C" int lvs_encoder_destroy(lvs_encoder* e) {
    lvs_lib_bind_thread();
    if (!e) return LVS_OK;
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->word); cudaFree(e->pos); cudaFree(e->type0); cudaFree(e->eln_w); cudaFree(e->eln_b); cudaFree(e->d_err);
    for (auto& L : e->layers) {
        for (Dense* d : {&L.qkv, &L.attn_out, &L.inter, &L.out}) { cudaFree(d->w); cudaFree(d->b); }
        cudaFree(L.ln1_w); cudaFree(L.ln1_b); cudaFree(L.ln2_w); cudaFree(L.ln2_b);
    }
    cudaFree(e->x); cudaFree(e->qkv); cudaFree(e->ctx); cudaFree(e->h); cudaFree(e->y); cudaFree(e->pooled); cudaFree(e->d_ids);
    if (e->stage) cudaFreeHost(e->stage);
    for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return LVS_OK;
}

static int stage_reserve(lvs_encoder* e, size_t bytes) {
    if (e->stage_bytes >= bytes) return LVS_OK;
    if (e->stage) cudaFreeHost(e->stage);
    e->stage = nullptr; e->stage_bytes = 0;
    LVS_CU(cudaHostAlloc(&e->stage, bytes, cudaHostAllocDefault));
    e->stage_bytes = bytes;
    return LVS_OK;
}

// fp32 host array -> device: as fp32 (dst_f32) or as bf16 rows at a column-major-free offset (dst_bf16 + row_off * in)
static int upload(lvs_encoder* e, const float* data, size_t n, float* dst_f32, __nv_bfloat16* dst_bf16) {
    const size_t chunk = (size_t)16 << 20;               // floats per staging pass
    int rc = stage_reserve(e, std::min(n, chunk) * 4);
    if (rc != LVS_OK) return rc;
    float* tmp = nullptr;
    if (dst_bf16) { rc = dev_alloc((void**)&tmp, std::min(n, chunk) * 4); if (rc != LVS_OK) return rc; }
    for (size_t o = 0; o < n; o += chunk) {
        const size_t m = std::min(chunk, n - o);
        memcpy(e->stage, data + o, m * 4);
        cudaError_t ce = cudaMemcpyAsync(dst_bf16 ? tmp : dst_f32 + o, e->stage, m * 4, cudaMemcpyHostToDevice, e->stream);
        if (ce == cudaSuccess && dst_bf16) { f32_to_bf16_kernel<<<1024, 256, 0, e->stream>>>(tmp, dst_bf16 + o, m); ce = cudaGetLastError(); }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) { cudaFree(tmp); return lvs_fail(LVS_ECUDA, "encoder: weight upload failed: %s", cudaGetErrorString(ce)); }
    }
    cudaFree(tmp);
    return LVS_OK;
}

extern "C" int lvs_encoder_load(lvs_encoder* e, const char* name, const float* data, int64_t n) {
    lvs_lib_bind_thread();
    if (!e || !name || !data) return lvs_fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(e->mu);
    const int H = e->cfg.hidden, I = e->cfg.intermediate;
    std::string s(name);
    for (const char* pre : {"roberta.", "model.", "encoder.roberta."}) if (s.rfind(pre, 0) == 0) { s = s.substr(strlen(pre)); break; }
    auto expect = [&](int64_t want) { return n == want ? LVS_OK : lvs_fail(LVS_EINVAL, "encoder: %s has %lld elements, expected %lld", name, (long long)n, (long long)want); };
    int rc = LVS_EINVAL;
    bool known = true;
    if (s == "embeddings.word_embeddings.weight") { if ((rc = expect((int64_t)e->cfg.vocab * H)) == LVS_OK) rc = upload(e, data, n, e->word, nullptr); }
    else if (s == "embeddings.position_embeddings.weight") { if ((rc = expect((int64_t)e->cfg.max_pos * H)) == LVS_OK) rc = upload(e, data, n, e->pos, nullptr); }
    else if (s == "embeddings.token_type_embeddings.weight") { if (n < H) rc = expect(H); else rc = upload(e, data, H, e->type0, nullptr); }   // row 0 is the one used
    else if (s == "embeddings.LayerNorm.weight") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, e->eln_w, nullptr); }
    else if (s == "embeddings.LayerNorm.bias") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, e->eln_b, nullptr); }
    else if (s.rfind("encoder.layer.", 0) == 0) {
        const size_t dot = s.find('.', 14);
        const int li = atoi(s.substr(14, dot - 14).c_str());
        if (dot == std::string::npos || li < 0 || li >= e->cfg.n_layers) return lvs_fail(LVS_EINVAL, "encoder: %s names a layer outside 0..%d", name, e->cfg.n_layers - 1);
        Layer& L = e->layers[li];
        const std::string t = s.substr(dot + 1);
        auto dense_w = [&](Dense& d, int row_off, int rows) { return (rc = expect((int64_t)rows * d.in)) == LVS_OK ? upload(e, data, n, nullptr, d.w + (size_t)row_off * d.in) : rc; };
        auto dense_b = [&](Dense& d, int off, int rows) { return (rc = expect(rows)) == LVS_OK ? upload(e, data, n, d.b + off, nullptr) : rc; };
        if (t == "attention.self.query.weight") rc = dense_w(L.qkv, 0, H);
        else if (t == "attention.self.key.weight") rc = dense_w(L.qkv, H, H);
        else if (t == "attention.self.value.weight") rc = dense_w(L.qkv, 2 * H, H);
        else if (t == "attention.self.query.bias") rc = dense_b(L.qkv, 0, H);
        else if (t == "attention.self.key.bias") rc = dense_b(L.qkv, H, H);
        else if (t == "attention.self.value.bias") rc = dense_b(L.qkv, 2 * H, H);
        else if (t == "attention.output.dense.weight") rc = dense_w(L.attn_out, 0, H);
        else if (t == "attention.output.dense.bias") rc = dense_b(L.attn_out, 0, H);
        else if (t == "attention.output.LayerNorm.weight") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, L.ln1_w, nullptr); }
        else if (t == "attention.output.LayerNorm.bias") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, L.ln1_b, nullptr); }
        else if (t == "intermediate.dense.weight") rc = dense_w(L.inter, 0, I);
        else if (t == "intermediate.dense.bias") rc = dense_b(L.inter, 0, I);
        else if (t == "output.dense.weight") rc = dense_w(L.out, 0, H);
        else if (t == "output.dense.bias") rc = dense_b(L.out, 0, H);
        else if (t == "output.LayerNorm.weight") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, L.ln2_w, nullptr); }
        else if (t == "output.LayerNorm.bias") { if ((rc = expect(H)) == LVS_OK) rc = upload(e, data, n, L.ln2_b, nullptr); }
        else known = false;
    } else known = false;
    if (!known) return lvs_fail(LVS_EINVAL, "encoder: %s is not a parameter of the encoder (pooler / position_ids buffers can be skipped by the caller)", name);
    if (rc == LVS_OK) { bool seen = false; for (auto& l : e->loaded) seen |= l == s; if (!seen) e->loaded.push_back(s); }
    return rc;
}

static int ws_reserve(lvs_encoder* e, int64_t tokens, int64_t seqs) {
    const int H = e->cfg.hidden, I = e->cfg.intermediate;
    int rc = LVS_OK;
    if (tokens > e->ws_tokens) {
        cudaFree(e->x); cudaFree(e->qkv); cudaFree(e->ctx); cudaFree(e->h); cudaFree(e->y); cudaFree(e->d_ids);
        e->x = e->qkv = e->ctx = e->h = nullptr; e->y = nullptr; e->d_ids = nullptr; e->ws_tokens = 0;
        auto A = [&](void** p, size_t bytes) { if (rc == LVS_OK) rc = dev_alloc(p, bytes); };
        A((void**)&e->x, (size_t)tokens * H * 2); A((void**)&e->qkv, (size_t)tokens * 3 * H * 2); A((void**)&e->ctx, (size_t)tokens * H * 2);
        A((void**)&e->h, (size_t)tokens * I * 2); A((void**)&e->y, (size_t)tokens * H * 4); A((void**)&e->d_ids, (size_t)tokens * 4);
        if (rc != LVS_OK) return rc;
        e->ws_tokens = tokens;
    }
    if (seqs > e->ws_seqs) {
        cudaFree(e->pooled); e->pooled = nullptr; e->ws_seqs = 0;
        if ((rc = dev_alloc((void**)&e->pooled, (size_t)seqs * H * 4)) != LVS_OK) return rc;
        e->ws_seqs = seqs;
    }
    return LVS_OK;
}

static int tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = lvs_lib_encode_tiled();
    if (!enc) return LVS_ECUDA;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kLinKC, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lvs_fail(LVS_ECUDA, "encoder: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LVS_OK;
}

template <int EPI>
static int launch_linear(lvs_encoder* e, const __nv_bfloat16* X, const Dense& d, int64_t M, const __nv_bfloat16* resid, void* out) {
    static std::once_flag once;
    static cudaError_t once_err = cudaSuccess;
    std::call_once(once, [] {
        once_err = cudaFuncSetAttribute(linear_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)linear_smem_bytes(false, EPI));
        if (once_err == cudaSuccess)
            once_err = cudaFuncSetAttribute(linear_kernel<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)linear_smem_bytes(true, EPI));
    });
    if (once_err != cudaSuccess) return lvs_fail(LVS_ECUDA, "encoder: cudaFuncSetAttribute failed: %s", cudaGetErrorString(once_err));
    // enough row blocks to keep every SM busy with 256-row tiles: pairs of CTAs share each weight tile (see linear_kernel.cuh)
    const int sm = lvs_lib_sm_count();
    const uint32_t tiles_n = (uint32_t)((d.out + kLinN - 1) / kLinN);
    const bool pair = e->opt_pair && (uint64_t)((M + 2 * kLinM - 1) / (2 * kLinM)) * tiles_n >= (uint64_t)(sm / 2);
    CUtensorMap tx, tw;
    int rc;
    if ((rc = tmap_2d(&tx, X, (uint64_t)M, (uint64_t)d.in, kLinM)) != LVS_OK) return rc;
    if ((rc = tmap_2d(&tw, d.w, (uint64_t)d.out, (uint64_t)d.in, pair ? kLinN / 2 : kLinN)) != LVS_OK) return rc;
    LinearParams p;
    p.M = (uint32_t)M; p.N = (uint32_t)d.out; p.K = (uint32_t)d.in;
    const int64_t bm = pair ? 2 * kLinM : kLinM;
    p.tiles_m = (uint32_t)((M + bm - 1) / bm); p.tiles_n = tiles_n;
    p.bias = d.b; p.resid = resid; p.out = out;
    if (pair) {
        const uint32_t walkers = std::min<uint32_t>(p.tiles_m * p.tiles_n, (uint32_t)(sm / 2));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(walkers * 2); cfg.blockDim = dim3(lin_threads(EPI)); cfg.dynamicSmemBytes = linear_smem_bytes(true, EPI); cfg.stream = e->stream;
        cudaLaunchAttribute la[1];
        la[0].id = cudaLaunchAttributeClusterDimension;
        la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
        cfg.attrs = la; cfg.numAttrs = 1;
        LVS_CU(cudaLaunchKernelEx(&cfg, linear_kernel<EPI, true>, tx, tw, p));
    } else {
        const uint32_t grid = std::min<uint32_t>(p.tiles_m * p.tiles_n, (uint32_t)sm);
        linear_kernel<EPI, false><<<grid, lin_threads(EPI), linear_smem_bytes(false, EPI), e->stream>>>(tx, tw, p);
        LVS_CU(cudaGetLastError());
    }
    return LVS_OK;
}

// ids are on the device (e->d_ids); leaves the pooled embeddings in e->pooled.  Enqueue only.
static int forward(lvs_encoder* e, int B, int L) {
    const lvs_encoder_config& c = e->cfg;
    const int H = c.hidden;
    const int64_t M = (int64_t)B * L;
    cudaStream_t st = e->stream;
    EmbedParams ep;
    ep.ids = e->d_ids; ep.B = B; ep.L = L; ep.H = H; ep.pad_id = c.pad_id; ep.max_pos = c.max_pos; ep.vocab = c.vocab;
    ep.word = e->word; ep.pos = e->pos; ep.type0 = e->type0; ep.ln_w = e->eln_w; ep.ln_b = e->eln_b; ep.eps = c.ln_eps; ep.out = e->x; ep.error = e->d_err;
    embed_ln_kernel<<<dim3((unsigned)B, (unsigned)((L + 63) / 64)), 256, (size_t)L * 4, st>>>(ep);
    LVS_CU(cudaGetLastError());
    const int Lp = (L + 63) / 64 * 64;
    const size_t asmem = attention_smem_bytes(Lp);
    if (Lp / 64 > kAttnMaxBlocks) return lvs_fail(LVS_ELIMIT, "encoder: sequences of %d tokens exceed the attention kernel's %d key blocks", L, kAttnMaxBlocks);
    if (asmem + 1024 > lvs_lib_smem_optin()) return lvs_fail(LVS_ELIMIT, "encoder: sequences of %d tokens do not fit the attention kernel's shared memory", L);
    {
        static std::once_flag once;
        static cudaError_t once_err = cudaSuccess;
        const int optin = (int)lvs_lib_smem_optin();
        std::call_once(once, [optin] {
            cudaFuncAttributes fa;                 // the kernel has a word of static shared memory: the opt-in limit covers both
            once_err = cudaFuncGetAttributes(&fa, attention_kernel);
            if (once_err == cudaSuccess)
                once_err = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
        });
        if (once_err != cudaSuccess) return lvs_fail(LVS_ECUDA, "encoder: cudaFuncSetAttribute failed: %s", cudaGetErrorString(once_err));
    }
    const unsigned ln_grid = (unsigned)std::min<int64_t>((M + 7) / 8, (int64_t)lvs_lib_sm_count() * 8);
    int rc;
    for (auto& Lr : e->layers) {
        if ((rc = launch_linear<EPI_BIAS>(e, e->x, Lr.qkv, M, nullptr, e->qkv)) != LVS_OK) return rc;
        attention_kernel<<<dim3((unsigned)((L + kAttnQB - 1) / kAttnQB), (unsigned)c.n_heads, (unsigned)B), 256, asmem, st>>>(e->qkv, e->d_ids, L, Lp, H, c.n_heads,
                                                                                                                       c.pad_id, e->ctx);
        LVS_CU(cudaGetLastError());
        if ((rc = launch_linear<EPI_BIAS_RESID>(e, e->ctx, Lr.attn_out, M, e->x, e->y)) != LVS_OK) return rc;
        add_ln_kernel<<<ln_grid, 256, 0, st>>>(e->y, (int)M, H, Lr.ln1_w, Lr.ln1_b, c.ln_eps, e->x);
        LVS_CU(cudaGetLastError());
        if ((rc = launch_linear<EPI_BIAS_GELU>(e, e->x, Lr.inter, M, nullptr, e->h)) != LVS_OK) return rc;
        if ((rc = launch_linear<EPI_BIAS_RESID>(e, e->h, Lr.out, M, e->x, e->y)) != LVS_OK) return rc;
        add_ln_kernel<<<ln_grid, 256, 0, st>>>(e->y, (int)M, H, Lr.ln2_w, Lr.ln2_b, c.ln_eps, e->x);
        LVS_CU(cudaGetLastError());
    }
    pool_kernel<<<dim3((unsigned)B, (unsigned)((H + 255) / 256)), 256, 0, st>>>(e->x, e->d_ids, L, H, c.pad_id, e->pooled);
    LVS_CU(cudaGetLastError());
    return LVS_OK;
}

static int run_forward(lvs_encoder* e, const int32_t* ids, int B, int L) {
    if (B < 1 || L < 1 || !ids) return lvs_fail(LVS_EINVAL, "encoder: bad ids / batch / length");
    if (L + e->cfg.pad_id + 1 > e->cfg.max_pos) return lvs_fail(LVS_ELIMIT, "encoder: %d tokens exceed the position table (%d)", L, e->cfg.max_pos);
    if (e->loaded.size() < e->n_expected) return lvs_fail(LVS_ESTATE, "encoder: %zu of %zu parameters are loaded", e->loaded.size(), e->n_expected);
    int rc;
    if ((rc = ws_reserve(e, (int64_t)B * L, B)) != LVS_OK) return rc;
    if ((rc = stage_reserve(e, (size_t)B * L * 4)) != LVS_OK) return rc;
    memcpy(e->stage, ids, (size_t)B * L * 4);
    LVS_CU(cudaMemcpyAsync(e->d_ids, e->stage, (size_t)B * L * 4, cudaMemcpyHostToDevice, e->stream));
    LVS_CU(cudaEventRecord(e->ev[0], e->stream));
    if ((rc = forward(e, B, L)) != LVS_OK) return rc;
    LVS_CU(cudaEventRecord(e->ev[1], e->stream));
    int32_t err = 0;
    LVS_CU(cudaMemcpyAsync(&err, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
    LVS_CU(cudaStreamSynchronize(e->stream));
    cudaEventElapsedTime(&e->last_ms, e->ev[0], e->ev[1]);
    if (err) { cudaMemset(e->d_err, 0, 4); return lvs_fail(LVS_EINVAL, "encoder: a token id is outside the vocabulary"); }
    return LVS_OK;
}

extern "C" int lvs_encoder_embed(lvs_encoder* e, const int32_t* ids, int B, int L, float* out) {
    lvs_lib_bind_thread();
    if (!e || !out) return lvs_fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = run_forward(e, ids, B, L);
    if (rc != LVS_OK) return rc;
    LVS_CU(cudaMemcpy(out, e->pooled, (size_t)B * e->cfg.hidden * 4, cudaMemcpyDeviceToHost));
    return LVS_OK;
}

extern "C" int lvs_encoder_embed_upsert(lvs_encoder* e, lvs_collection* c, const int32_t* ids, int B, int L, const int64_t* rows,
                                        const uint32_t* codes, const uint64_t* ties) {
    lvs_lib_bind_thread();
    if (!e || !c) return lvs_fail(LVS_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lk(e->mu);
    int rc = run_forward(e, ids, B, L);          // synchronised: the pooled vectors are complete in HBM
    if (rc != LVS_OK) return rc;
    return lvs_upsert_device_vectors(c, e->pooled, LVS_DT_F32, B, rows, codes, ties);
}

extern "C" int lvs_encoder_last_ms(const lvs_encoder* e, float* ms) {
    if (!e || !ms) return lvs_fail(LVS_EINVAL, "NULL argument");
    *ms = e->last_ms;
    return LVS_OK;
}
