// C-ABI library (include/sfq_b200.h): host orchestration of the sm_100a kernels.
// No CPU coding path exists in this file: every stream byte is produced by a kernel launch.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <chrono>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfq_b200.h"
#include "sfq_kernels.cuh"
#include "sfq_layout.h"
#include "sfq_worm.h"

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool view = false;                      // a sub-range of ctx->scratch: not owned
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && !view) return cudaSuccess;
        if (p && !view) cudaFree(p);
        p = nullptr; cap = 0; view = false;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; view = false; }
    // the next `bytes` of `pool` from *off (256-byte aligned); the caller has sized the pool
    void carve(const DevBuf &pool, size_t bytes, size_t *off) {
        if (p && !view) cudaFree(p);
        p = static_cast<uint8_t *>(pool.p) + *off; cap = bytes; view = true;
        *off += (bytes + 255) & ~(size_t)255;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum { EV_START = 0, EV_H2D, EV_SCAN, EV_PLAN, EV_CODE_END, EV_D2H, EV_COUNT };
enum { WEV = 10 };   // events per wave

}  // namespace

struct sfq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};     // the qlt and rec coder kernels run beside gen
    cudaEvent_t fork_ev = nullptr, join_ev[2] = {nullptr, nullptr};
    std::string err;
    sfq_stats st{};
    uint32_t max_resident = 0;
    uint64_t chunk_phase = 0;               // sfq_set_chunk_phase: grid offset of the next compress call's input
    uint32_t lanes = 0;                     // chunk-streams per gen/rec coder warp (SFQ_LANES); 0 = by wave size
    int sm_count = 148;
    uint32_t rc_lanes = 8;                  // chunk-streams per warp of the coder-chain kernel (SFQ_RC_LANES)
    cudaEvent_t ev[EV_COUNT]{};
    std::vector<cudaEvent_t> wave_ev;       // 10 per wave: clear start, code start, code end, pack end, then start/end of gen, qlt, rec
    // device buffers (grow-only, reused across calls)
    // One arena for everything a call needs only while it runs and that scales with the resident chunks or
    // the input: compress carves its stream arenas and coding-step arrays from it, decompress its planes,
    // per-record tables and quality tables.  Alternating calls reuse the same memory instead of each keeping
    // a private set alive (which would halve the chunks that fit in one wave).
    DevBuf scratch;
    DevBuf text, out, tiles, tile_prefix, lines, scalars, rec_begin, r0, r1, metas, arenas, arena_buf,
           blob_off, gtab, qtab, pw, dchunks, bhdrs, bases, quals, hdrs, rec_chunk,
           t_llen, t_qlen, t_hlen, t_pfg, t_pfq, t_boff, t_qoff, t_hoff, t_ooff,
           e2_gsteps, e2_qkey, e2_qb, e2_sorted, e2_qsteps, e2_cnt, e2_esorted, e2_esteps, e2_segs, e2_ctr, e2_chunks, rec_qoff,
           e2_gbins, e2_gcnt, rec_boff;
    int spread = -1;                        // SFQ_SPREAD=0..3 (default: 3 for large waves, else 0; see decompress_on_device)
    unsigned spread_smem[3] = {0, 0, 0};    // dynamic shared memory reserved per CTA: base, quality, header decoder
    uint32_t dec_warps = 4;                 // SFQ_DEC_WARPS=1..4: warps per CTA of the thread-per-chunk decoders (one-warp CTAs each
                                            // carry the whole static shared memory: 4x the footprint, less L1 for everyone - 35 % slower decode)
    int enc_order = 0;                      // SFQ_ENC_ORDER=1: quality path's keys+scan before k_gen_model
    uint32_t enc_rec_lanes = 0;             // SFQ_ENC_REC_LANES: chunk-streams per warp of the header encoder (0 = pick_lanes)
    cudaEvent_t head_ev = nullptr;
    cudaEvent_t gbins_ev = nullptr;         // k_gen_replay has read the partition lists: their memory may become the quality steps
    bool qch = false;                       // SFQ_QCH=1: the 4-lane quality decoder with the compact-header entry (7 sectors read, 2 dirtied per quality instead of 8 / 5;
                                            // A/B at 10 GB: DRAM traffic of the kernel -30 %, but five more loads and three selects per link: 943 ms against 915)
    uint32_t rc_warps = 1;                  // SFQ_RC_WARPS=1..4: warps per CTA of the base coder chain sharing one reciprocal table (A/B r2af: 4 frees 24 KB of
                                            // shared memory per SM and changes nothing: 576 against 571 ms per wave)
    bool rc_q4 = true;                      // SFQ_RC_Q4=0: the quality coder chain with one lane per chunk-stream (k_rc_encode<1>) instead of four
    uint32_t dec_hold_us = 0;               // SFQ_DEC_HOLD_US: device-side delay in front of the base decoder of a large wave (A/B r2ah: with 100 us the base
                                            // decoder's CTAs are packed two to an SM in 60 of 60 calls - the even spread needs the kernels to arrive TOGETHER)
    bool rec_global = false;                // SFQ_REC_GLOBAL=1 (see the header coder's launch)
    DevBuf rec_scr;
    int dec_sched = 0;                      // SFQ_DEC_SCHED=1: header decoder after the base decoder instead of beside it
    int enc_sched = 0;                      // SFQ_ENC_SCHED: when the header encoder / base coder chain of a wave may start (see compress_on_device)
    cudaEvent_t qp2_ev = nullptr, qtp_ev = nullptr;   // quality path: second-level partition done / model replay done
    bool marks = false;                     // SFQ_MARKS=1: host-side time marks of a call on stderr (without the per-kernel events of SFQ_TRACE)
    bool alias_steps = true;                // SFQ_ALIAS=0: partition lists (gbins) and quality steps (qsteps) in separate memory, as before
    cudaStream_t copy_in = nullptr, copy_out = nullptr;          // sfq_compress of a large host buffer: parts copied in / out beside the coding
    cudaEvent_t part_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // [0,1] text buffer filled, [2] part coded, [3] last copy out done
    uint32_t dec_lanes = 0;                 // SFQ_DEC_LANES: chunk-streams per warp of the thread-per-chunk decoders (0 = pick_lanes / dec_fit)
    bool dec_fit = false;                   // SFQ_DEC_FIT=1: widen the warps of a large wave so that its CTAs number at most one per SM (A/B: 17 lanes per warp cost the base decoder 3 %, and no outlier either way in 6 steps)
    double head_frac = 0.08;                // SFQ_HEAD_FRAC: size of the head part of a pipelined sfq_compress, as a fraction of a coder wave (swept 0.05 .. 0.25 at 10 GB: 815 / 821 / 827 / 841 ms per call)
    int parts = 0;                          // SFQ_PARTS: parts per sfq_compress call (0 = a head part + one per coder wave the input needs, from 256 MB; -1 = that for any size)
    bool trace = false;                     // SFQ_TRACE=1: per-kernel event timings of the coder waves on stderr
    std::vector<std::pair<const char *, std::pair<cudaEvent_t, cudaEvent_t>>> tr;
    int gdec32 = 0;                         // SFQ_GDEC=1: warp-converged base decoder (A/B; slower)
    uint32_t qlpc = 0;                      // SFQ_QLPC=4|8: lanes per chunk of the quality decoder (0 = by wave size)
    int gm_variant = 0;                     // SFQ_GM_VARIANT: register budget / batch of k_gen_model (A/B runs)
    bool dec_gen_first = false;             // SFQ_DEC_ORDER=1: base decoder launched first, on the high-priority stream (A/B: consistently packed, 2x slower)
    bool qspec = false;                     // SFQ_QSPEC=1: the quality decoder prefetches the model of the "same symbol again" context (A/B: 4 % faster alone, 2 % slower beside the other decoders)
    bool enc_prio_gen = true;               // SFQ_ENC_PRIO=0: the high-priority stream goes to the quality path of a compress wave
    int plane_mask = 7;                     // per-plane test hooks: 1 = gen, 2 = qlt, 4 = rec paths run (7 = the product)
    bool q_scatter1 = false;                // SFQ_QSCATTER=1: one-pass quality scatter over 65 536 global cursors (the round-1 form)
    bool gm_table = false;                  // SFQ_GM_TABLE=1: base models in a global hash table (k_gen_model, the round-1 form) instead of partitioned replay
    int gen_ahead2 = -1;                    // SFQ_GEN_AHEAD2=0/1: base decoder's two-ahead line prefetch (default on)
    bool qdec_octets = true;                // SFQ_QDEC=0: the first (sub-warp mask) quality decoder, for A/B runs
    bool serial_roles = false;              // SFQ_SERIAL_ROLES=1: gen, qlt, rec kernels of a wave one after another (diagnosis)
    uint32_t qgpw = 2;                      // quality-decoder groups (chunks) per warp (SFQ_QGPW: 1, 2 or 4)
    bool serial_encoder = false;            // SFQ_ENC_SERIAL=1: single-pass coders (one chain per chunk-stream) for A/B runs
    HostBuf h_out, h_out_d, h_small;        // results of compress / decompress (separate: a container returned by
                                            // sfq_compress may be handed straight to sfq_decompress)
    void release_all() {
        DevBuf *all[] = {&text, &out, &tiles, &tile_prefix, &lines, &scalars, &rec_begin, &r0, &r1, &metas,
                         &arenas, &arena_buf, &blob_off, &gtab, &qtab, &pw, &dchunks, &bhdrs, &bases, &quals,
                         &hdrs, &rec_chunk, &t_llen, &t_qlen, &t_hlen, &t_pfg, &t_pfq, &t_boff, &t_qoff,
                         &t_hoff, &t_ooff, &rec_scr, &e2_gsteps, &e2_qkey, &e2_qb, &e2_sorted, &e2_qsteps, &e2_cnt, &e2_esorted,
                         &e2_esteps, &e2_segs, &e2_ctr, &e2_chunks, &rec_qoff, &e2_gbins, &e2_gcnt, &rec_boff};
        for (DevBuf *b : all) b->release();
        scratch.release();
        h_out.release(); h_out_d.release(); h_small.release();
    }
};

namespace {

void release_workspace(sfq_ctx *ctx) {
    DevBuf *views[] = {&ctx->arena_buf, &ctx->qtab, &ctx->bases, &ctx->quals, &ctx->hdrs, &ctx->t_llen, &ctx->t_qlen, &ctx->t_hlen,
                       &ctx->t_pfg, &ctx->t_pfq, &ctx->t_boff, &ctx->t_qoff, &ctx->t_hoff, &ctx->t_ooff,
                       &ctx->e2_gsteps, &ctx->e2_qkey, &ctx->e2_qb, &ctx->e2_sorted, &ctx->e2_qsteps, &ctx->e2_cnt,
                       &ctx->e2_esorted, &ctx->e2_esteps, &ctx->e2_segs, &ctx->e2_gbins, &ctx->e2_gcnt};
    for (DevBuf *d : views) d->release();
    ctx->gtab.release(); ctx->pw.release();
    ctx->scratch.release();
}
// Large caller-facing buffers win over the cached coder workspace: on OOM drop it and retry.
cudaError_t ensure_big(sfq_ctx *ctx, DevBuf &b, size_t bytes) {
    cudaError_t e = b.ensure(bytes);
    if (e != cudaErrorMemoryAllocation) return e;
    cudaGetLastError();
    release_workspace(ctx);
    return b.ensure(bytes);
}

int fail(sfq_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SFQ_ERR_NOMEM : SFQ_ERR_CUDA,        \
                        "CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);      \
    } while (0)
#define LAUNCHED() (ctx->st.kernel_launches++)

// The reference's croak texts (config.cpp:54-68 prefix omitted), keyed by kernel status.
int status_to_error(sfq_ctx *ctx, const SfqChunkMeta &m, uint64_t chunk, uint64_t rec_base) {
    const unsigned long long rec = rec_base + m.status_arg;
    switch (m.status) {
    case SFQ_E_AT: return fail(ctx, SFQ_ERR_FASTQ, "fastq file: expecting '@' at record %llu (chunk %llu)", rec, (unsigned long long)chunk);
    case SFQ_E_PLUS: return fail(ctx, SFQ_ERR_FASTQ, "fastq file: expecting '+' at record %llu (chunk %llu)", rec, (unsigned long long)chunk);
    case SFQ_E_TRUNC: return fail(ctx, SFQ_ERR_FASTQ, "fastq file: record seems truncated  after record %llu", rec);
    case SFQ_E_OVERSIZE: return fail(ctx, SFQ_ERR_FASTQ, "wierd second id at record %llu", rec);
    case SFQ_E_BASE: return fail(ctx, SFQ_ERR_FASTQ, "unexpected genome char: %c", (int)(m.status_arg & 0xffu));
    case SFQ_E_NBYTE: return fail(ctx, SFQ_ERR_FASTQ, "switched N_byte: %c", (int)m.status_arg);
    case SFQ_E_SEPS: return fail(ctx, SFQ_ERR_FASTQ, "ERROR: irregulal record (over 64 non alpha non digit). Is it a valid fastq file?");
    case SFQ_E_FIRSTHDR: return fail(ctx, SFQ_ERR_UNSUPPORTED, "chunk %llu: first header longer than 399 chars", (unsigned long long)chunk);
    case SFQ_E_EMPTYSEQ: return fail(ctx, SFQ_ERR_UNSUPPORTED, "chunk %llu: first record has an empty base line", (unsigned long long)chunk);
    case SFQ_E_CHUNKSIZE: return fail(ctx, SFQ_ERR_UNSUPPORTED, "chunk %llu holds 4 GiB or more of bases, qualities or headers: use a smaller chunk size (-c; -R codes the file as one chunk)", (unsigned long long)chunk);
    case SFQ_E_CORRUPT: return fail(ctx, SFQ_ERR_FORMAT, "chunk %llu: corrupt stream", (unsigned long long)chunk);
    default: return fail(ctx, SFQ_ERR_CUDA, "chunk %llu: internal status %u", (unsigned long long)chunk, m.status);
    }
}

// How many chunks can have their model tables resident at once.
uint32_t pick_resident(sfq_ctx *ctx, uint64_t nchunks, uint64_t per_chunk, uint64_t already_have, uint64_t fixed = 0) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    uint64_t budget = (uint64_t)((double)(free_b + already_have) * 0.92);
    budget = budget > fixed ? budget - fixed : 0;
    uint64_t r = budget / per_chunk;
    if (r < 1) r = 1;
    if (ctx->max_resident && r > ctx->max_resident) r = ctx->max_resident;
    const char *env = getenv("SFQ_MAX_RESIDENT");
    if (env && atoll(env) > 0 && r > (uint64_t)atoll(env)) r = (uint64_t)atoll(env);
    if (r > nchunks) r = nchunks;
    // even waves: ceil(nchunks / nwaves)
    uint64_t nw = (nchunks + r - 1) / r;
    r = (nchunks + nw - 1) / nw;
    return (uint32_t)r;
}

// Chunk-streams per warp of the thread-per-chunk coders: few when the wave is small (the chains are then
// latency-bound and a lane's slow path stalls fewer neighbours), more when it is large (issue slots are
// then the scarce resource and a fuller warp spends fewer of them per symbol).
uint32_t pick_lanes(const sfq_ctx *ctx, uint32_t nc) { return ctx->lanes ? ctx->lanes : nc >= 4096u ? 16u : nc >= 2048u ? 8u : 4u; }

// kernels index the per-chunk pools by the chunk's position in the wave
SfqWorkspace ws_at(const SfqWorkspace &ws, uint32_t) { return ws; }

float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
// SFQ_TRACE: host wall-clock marks (where the host thread waits or works while the device could be idle)
void host_mark(const sfq_ctx *ctx, const char *what);

// SFQ_TRACE: bracket a launch with events on its stream; trace_dump() prints and frees them after the sync.
struct TraceScope {
    sfq_ctx *ctx; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; const char *name;
    TraceScope(sfq_ctx *c, const char *n, cudaStream_t s) : ctx(c), st(s), name(n) {
        if (ctx->trace) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~TraceScope() { if (ctx->trace) { cudaEventRecord(b, st); ctx->tr.push_back({name, {a, b}}); } }
};
void trace_dump(sfq_ctx *ctx, const char *what) {
    if (!ctx->trace) return;
    cudaEvent_t t0 = ctx->tr.empty() ? nullptr : ctx->tr[0].second.first;
    for (auto &e : ctx->tr) {
        float off = 0, ms = 0;
        cudaEventElapsedTime(&off, t0, e.second.first); cudaEventElapsedTime(&ms, e.second.first, e.second.second);
        fprintf(stderr, "[sfq trace] %s %-22s start %9.3f ms  dur %9.3f ms\n", what, e.first, off, ms);
        if (e.second.first != t0) cudaEventDestroy(e.second.first);
        cudaEventDestroy(e.second.second);
    }
    if (t0) cudaEventDestroy(t0);
    ctx->tr.clear();
}
#define TRACED(name, stream, launch) { TraceScope ts_(ctx, name, stream); launch; }

int ensure_wave_events(sfq_ctx *ctx, size_t waves) {
    while (ctx->wave_ev.size() < waves * WEV) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        ctx->wave_ev.push_back(e);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------ compress
// A part of a file coded as the continuation of a container under construction (sfq_compress pipelines the parts of a large
// host buffer: copy in, code, copy out): blobs go to d_out from `cursor` on, the index entries are collected by the caller,
// no file header and no index are written.
struct PartIo {
    uint64_t cursor;                     // in: where this part's blobs start in d_out; out: where they end
    std::vector<uint64_t> *index;        // blob offsets of all parts so far
    uint64_t out_total;                  // bytes the parts so far decode to
    // Called once the part's coder waves are enqueued: the next part's text is sent off only now.  Sent earlier, the 5 GB
    // transfer would sit in the host-to-device copy engine ahead of this part's own small set-up copies (arena layout,
    // chunk tables: pageable, hence synchronous), and the waves would start ~90 ms late behind it (measured).
    std::function<int()> on_enqueued;
};
int compress_on_device(sfq_ctx *ctx, const uint8_t *d_text, size_t n, int level, uint64_t chunk_bytes,
                       uint8_t *d_out, size_t out_cap, size_t *out_n, PartIo *part = nullptr) {
    cudaStream_t s = ctx->stream;
    cudaStream_t side0 = ctx->serial_roles ? s : ctx->side[0], side1 = ctx->serial_roles ? s : ctx->side[1];
    sfq_stats &st = ctx->st;
    level = level > 4 ? 4 : level < 1 ? 1 : level;                    // range_level, config.cpp:231-236
    const uint64_t phase_arg = ctx->chunk_phase;                      // consumed by this call whatever its outcome
    ctx->chunk_phase = 0;
    if (!chunk_bytes) chunk_bytes = 1ull << 20;
    if (chunk_bytes < 4096) chunk_bytes = 4096;
    if (n == 0) return fail(ctx, SFQ_ERR_FASTQ, "no records were found");
    if (((uintptr_t)d_text & 15) || ((uintptr_t)d_out & 15)) return fail(ctx, SFQ_ERR_ARG, "device buffers must be 16-byte aligned");
    if (out_cap < sizeof(SfqFileHeader)) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small");

    // ---- scan: newline index
    const uint64_t ntiles = (n + SFQ_SCAN_TILE - 1) / SFQ_SCAN_TILE;
    CK(ctx->tiles.ensure(ntiles * 4));
    CK(ctx->tile_prefix.ensure(ntiles * 8));
    CK(ctx->scalars.ensure(256));
    CK(ctx->h_small.ensure(4096));
    uint64_t *d_scal = ctx->scalars.as<uint64_t>();     // [0] newline total, [1] blob cursor, [2] pack overflow
    k_count_newlines<<<(unsigned)ntiles, SFQ_SCAN_THREADS, 0, s>>>(d_text, n, ctx->tiles.as<uint32_t>()); LAUNCHED();
    k_scan_tiles<<<1, 1024, 0, s>>>(ctx->tiles.as<uint32_t>(), ctx->tile_prefix.as<uint64_t>(), ntiles, d_scal); LAUNCHED();
    uint64_t *h_small = ctx->h_small.as<uint64_t>();
    CK(cudaMemcpyAsync(h_small, d_scal, 8, cudaMemcpyDeviceToHost, s));
    uint8_t last_byte = 0;
    CK(cudaMemcpyAsync(&h_small[1], d_text + n - 1, 1, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    host_mark(ctx, "compress: newline count known");
    const uint64_t nlines = h_small[0];
    last_byte = *reinterpret_cast<uint8_t *>(&h_small[1]);
    if (last_byte != '\n' || nlines % 4 || nlines == 0)
        return fail(ctx, SFQ_ERR_FASTQ, "fastq file: record seems truncated  after record %llu", (unsigned long long)(nlines / 4));
    const uint64_t nrec_total = nlines / 4;
    CK(ctx->lines.ensure((nlines + 1) * 8));
    const uint64_t *d_ls = ctx->lines.as<uint64_t>();
    k_fill_lines<<<(unsigned)ntiles, SFQ_SCAN_THREADS, 0, s>>>(d_text, n, ctx->tile_prefix.as<uint64_t>(), ctx->lines.as<uint64_t>()); LAUNCHED();
    CK(cudaEventRecord(ctx->ev[EV_SCAN], s));

    // ---- plan: chunk boundaries, then per-chunk framing facts
    const uint64_t phase = phase_arg;                         // (one call only: sfq_set_chunk_phase)
    const uint64_t nslots = sfq_slot_count(n, chunk_bytes, phase);
    CK(ctx->rec_begin.ensure((nslots + 1) * 8));
    k_chunk_bounds<<<(unsigned)((nslots + 1 + 127) / 128), 128, 0, s>>>(d_ls, nrec_total, chunk_bytes, nslots, phase, ctx->rec_begin.as<uint64_t>()); LAUNCHED();
    std::vector<uint64_t> rb(nslots + 1);
    CK(cudaMemcpyAsync(rb.data(), ctx->rec_begin.p, (nslots + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<uint64_t> r0, r1;
    for (uint64_t c = 0; c < nslots; c++)
        if (rb[c + 1] > rb[c]) { r0.push_back(rb[c]); r1.push_back(rb[c + 1]); }
    const uint32_t nchunks = (uint32_t)r0.size();
    CK(ctx->r0.ensure(nchunks * 8ull)); CK(ctx->r1.ensure(nchunks * 8ull));
    CK(ctx->metas.ensure(nchunks * sizeof(SfqChunkMeta)));
    CK(cudaMemcpyAsync(ctx->r0.p, r0.data(), nchunks * 8ull, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->r1.p, r1.data(), nchunks * 8ull, cudaMemcpyHostToDevice, s));
    SfqChunkMeta *d_metas = ctx->metas.as<SfqChunkMeta>();
    CK(ctx->rec_qoff.ensure(nrec_total * 4 + 64)); CK(ctx->rec_boff.ensure(nrec_total * 4 + 64));
    k_chunk_plan<<<(nchunks + 63) / 64, 64, 0, s>>>(d_text, d_ls, ctx->r0.as<uint64_t>(), ctx->r1.as<uint64_t>(), d_metas, nchunks, ctx->rec_qoff.as<uint32_t>(), ctx->rec_boff.as<uint32_t>()); LAUNCHED();
    std::vector<SfqChunkMeta> metas(nchunks);
    CK(cudaMemcpyAsync(metas.data(), d_metas, nchunks * sizeof(SfqChunkMeta), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[EV_PLAN], s));
    CK(cudaStreamSynchronize(s));
    host_mark(ctx, "compress: chunk plan known");
    uint64_t max_bases = 0, out_total = 0;
    st.nchunks = nchunks; st.nrecords = nrec_total; st.nbases = st.nquals = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        if (metas[c].status) return status_to_error(ctx, metas[c], c, r0[c]);
        max_bases = std::max<uint64_t>(max_bases, metas[c].nbases);
        st.nbases += metas[c].nbases; st.nquals += metas[c].nquals;
        out_total += metas[c].out_len;
    }

    // ---- code + pack, in waves of resident chunks; rerun with more room if a stream or table overflowed
    CK(ctx->arenas.ensure(nchunks * sizeof(SfqArena)));
    CK(ctx->blob_off.ensure(nchunks * 8ull));
    SfqArena *d_arenas = ctx->arenas.as<SfqArena>();
    uint64_t *d_blob_off = ctx->blob_off.as<uint64_t>();
    std::vector<SfqArena> arenas(nchunks);
    std::vector<uint64_t> blob_off(nchunks);
    uint64_t end_cursor = 0;
    st.retries = 0;
    for (uint32_t grow = 0;; grow++) {
        uint64_t max_nb = 0, max_nq = 0;
        for (uint32_t c = 0; c < nchunks; c++) { max_nb = std::max<uint64_t>(max_nb, metas[c].nbases); max_nq = std::max<uint64_t>(max_nq, metas[c].nquals); }
        // The two-phase encoder packs stream positions into 24 bits: a chunk of 2^24 or more bases or qualities (CLI -R on a
        // large file, a huge -c) goes through the single-pass coders instead (one chain per chunk-stream: slow, any size).
        const bool two_phase = !ctx->serial_encoder && max_nq < (1u << 24) && max_nb < (1u << 24);
        const bool gm_table = !two_phase || ctx->gm_table;            // base models in a global table (single-pass coder, or A/B)
        const uint32_t hbits = sfq_gen_hbits(level, max_bases, grow);
        const uint32_t cbits = sfq_q_cbits(level, grow);
        const uint32_t gp_bits = sfq_gen_gp_bits(max_nb, grow);
        const uint64_t gstride = gm_table ? sfq_gtable_bytes(level, hbits) : 0, qbytes = sfq_qhash_bytes(level, cbits), pbytes = sfq_pwpool_bytes();
        uint64_t max_arena = 0;
        for (uint32_t c = 0; c < nchunks; c++) {
            uint64_t end; SfqArena a;
            sfq_arena_layout(&metas[c], grow, 0, &a, &end);
            max_arena = std::max(max_arena, end);
        }
        auto esc_cap = [grow](uint64_t nq) { return std::min<uint64_t>(nq, (nq >> 4) << grow) + 64; };
        auto seg_cap = [](uint64_t nq) { return std::min<uint64_t>(SFQ_Q_NCTX, nq) + 1; };
        // two-phase encoder: no quality-model table in global memory, but the coding steps of every symbol
        // The base path's partition lists (8 B per base, dead once k_gen_replay has run) and the quality path's coding steps
        // (8 B per quality, first written by k_qlt_part1) share one range of the arena: the quality path waits for gbins_ev
        // before it touches it.  That is a quarter of a chunk's workspace, and what lets the 10 GB workload code in ONE wave.
        const bool alias = two_phase && !gm_table && ctx->alias_steps;
        const uint64_t e2_per_chunk = 4 * max_nb + (gm_table ? 0 : (alias ? 8 * (std::max(max_nb, max_nq) - max_nq) : 8 * max_nb) + (36ull << gp_bits)) + 15 * max_nq + SFQ_Q_CNT * 4ull + 16 * seg_cap(max_nq) + 12 * esc_cap(max_nq) + 256;
        const uint64_t per_chunk = gstride + pbytes + max_arena + 4096 + (two_phase ? e2_per_chunk : qbytes);
        const uint64_t have_now = ctx->scratch.cap;
        const uint32_t R = pick_resident(ctx, nchunks, per_chunk, have_now);
        const uint32_t nwaves = (nchunks + R - 1) / R;
        if (ensure_wave_events(ctx, nwaves)) return SFQ_ERR_CUDA;
        st.waves = nwaves; st.resident_chunks = R; st.workspace_bytes = (uint64_t)R * per_chunk;
        // arena layout: offsets restart at 0 for every wave; same for the coding-step arrays
        uint64_t wave_arena_max = 0, wave_nb = 0, wave_nq = 0, wave_ne = 0, wave_seg = 0;
        std::vector<SfqEnc2Chunk> e2c(nchunks);
        for (uint32_t w = 0; w < nwaves; w++) {
            uint64_t o = 0, gb = 0, qb = 0, eb = 0, sg = 0;
            for (uint32_t c = w * R; c < std::min(nchunks, (w + 1) * R); c++) {
                sfq_arena_layout(&metas[c], grow, o, &arenas[c], &o);
                e2c[c].goff = gb; e2c[c].qoff = qb; e2c[c].eoff = eb; e2c[c].ecap = (uint32_t)esc_cap(metas[c].nquals); e2c[c].pad = 0;
                gb += (metas[c].nbases + 3ull) & ~3ull; qb += (metas[c].nquals + 15ull) & ~15ull; eb += e2c[c].ecap; sg += seg_cap(metas[c].nquals);
            }
            wave_arena_max = std::max(wave_arena_max, o);
            wave_nb = std::max(wave_nb, gb); wave_nq = std::max(wave_nq, qb); wave_ne = std::max(wave_ne, eb); wave_seg = std::max(wave_seg, sg);
        }
        SfqEnc2Ws e2{};
        {   // everything wave-sized comes out of the shared scratch arena
            size_t sz[] = {wave_arena_max + 64, wave_nb * 4 + 256, wave_nq * 2 + 64, wave_nq + 64, wave_nq * 4 + 64, wave_nq * 8 + 256,
                                 (size_t)R * SFQ_Q_CNT * 4, wave_ne * 4 + 64, wave_ne * 8 + 64, wave_seg * sizeof(SfqSeg),
                                 gm_table ? 0 : (wave_nb + ((size_t)R * 4 << gp_bits)) * 8 + 256, gm_table ? 0 : ((size_t)R * 4 << gp_bits) + 64, (size_t)R * qbytes,
                                 gm_table ? (size_t)R * gstride : 0, (size_t)R * pbytes};
            if (alias) { sz[5] = std::max(sz[5], sz[10]); sz[10] = 0; }
            DevBuf *bufs[] = {&ctx->arena_buf, &ctx->e2_gsteps, &ctx->e2_qkey, &ctx->e2_qb, &ctx->e2_sorted, &ctx->e2_qsteps,
                              &ctx->e2_cnt, &ctx->e2_esorted, &ctx->e2_esteps, &ctx->e2_segs, &ctx->e2_gbins, &ctx->e2_gcnt, &ctx->qtab,
                              &ctx->gtab, &ctx->pw};
            // (the base-model table of the single-pass coder and the 256-symbol models come out of the same arena, so that
            // compress and decompress calls on one context reuse one allocation)
            auto used = [&](int k) { return k >= 13 || (two_phase ? k < 12 : (k == 0 || k == 12)); };
            size_t total = 0;
            for (int k = 0; k < 15; k++) if (used(k)) total += (sz[k] + 255) & ~(size_t)255;
            {
                cudaError_t e = ctx->scratch.ensure(total);
                if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); release_workspace(ctx); e = ctx->scratch.ensure(total); }
                CK(e);
            }
            size_t off = 0;
            for (int k = 0; k < 15; k++) if (used(k)) bufs[k]->carve(ctx->scratch, sz[k], &off);
            if (alias) { size_t o = static_cast<uint8_t *>(ctx->e2_qsteps.p) - static_cast<uint8_t *>(ctx->scratch.p); ctx->e2_gbins.carve(ctx->scratch, sz[5], &o); }
        }
        if (two_phase) {
            CK(ctx->e2_ctr.ensure(64)); CK(ctx->e2_chunks.ensure(nchunks * sizeof(SfqEnc2Chunk)));
            CK(cudaMemcpyAsync(ctx->e2_chunks.p, e2c.data(), nchunks * sizeof(SfqEnc2Chunk), cudaMemcpyHostToDevice, s));
            e2.gsteps = ctx->e2_gsteps.as<uint32_t>(); e2.qkey = ctx->e2_qkey.as<uint16_t>(); e2.qb = ctx->e2_qb.as<uint8_t>();
            e2.sorted = ctx->e2_sorted.as<uint32_t>(); e2.qsteps = ctx->e2_qsteps.as<uint64_t>(); e2.cnt = ctx->e2_cnt.as<uint32_t>();
            e2.esorted = ctx->e2_esorted.as<uint32_t>(); e2.esteps = ctx->e2_esteps.as<uint64_t>(); e2.segs = ctx->e2_segs.as<SfqSeg>();
            e2.seg_cap = wave_seg; e2.ctr = ctx->e2_ctr.as<uint32_t>();
            e2.gbins = ctx->e2_gbins.as<SfqU2>(); e2.gcnt = ctx->e2_gcnt.as<uint32_t>(); e2.gp_bits = gp_bits;
        }
        host_mark(ctx, "compress: workspace carved");
        CK(cudaMemcpyAsync(d_arenas, arenas.data(), nchunks * sizeof(SfqArena), cudaMemcpyHostToDevice, s));
        h_small[0] = 0; h_small[1] = part ? part->cursor : sizeof(SfqFileHeader); h_small[2] = 0;
        CK(cudaMemcpyAsync(d_scal, h_small, 24, cudaMemcpyHostToDevice, s));
        SfqWorkspace ws{};
        ws.gtab = ctx->gtab.as<uint8_t>(); ws.gtab_stride = gstride; ws.hbits = hbits;
        ws.qtab = ctx->qtab.as<uint32_t>(); ws.qtab_words = qbytes / 4; ws.cbits = level <= 1 ? 4096u : 1u << cbits; ws.pw = ctx->pw.as<uint32_t>();
        for (uint32_t w = 0; w < nwaves; w++) {
            const uint32_t c0 = w * R, nc = std::min(nchunks, c0 + R) - c0;
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 0], s));
            if (gm_table) CK(cudaMemsetAsync(ctx->gtab.p, 0, nc * gstride, s));
            else CK(cudaMemsetAsync(ctx->e2_gcnt.p, 0, (size_t)nc * 4 << gp_bits, s));
            if (!two_phase) CK(cudaMemsetAsync(ctx->qtab.p, 0, nc * qbytes, s));
            else { CK(cudaMemsetAsync(ctx->e2_cnt.p, 0, (uint64_t)nc * SFQ_Q_CNT * 4, s)); CK(cudaMemsetAsync(ctx->e2_ctr.p, 0, 64, s)); }
            CK(cudaMemsetAsync(ctx->pw.p, 0, nc * pbytes, s));
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 1], s));
            {   // fork: three paths side by side, join before packing.  The base path is the longest of a wave since the quality path's
                // scatter became cheap, so it gets the high-priority stream (its shared-memory-hungry CTAs are placed first);
                // SFQ_ENC_PRIO=0 gives that stream to the quality path again, as in round 1
                cudaStream_t sg = ctx->enc_prio_gen ? side0 : s, sq = ctx->enc_prio_gen ? s : side0;
                const uint32_t lanes = pick_lanes(ctx, nc);
                const unsigned nb = (nc + lanes - 1) / lanes;
                const unsigned nwarp_blocks = (nc * 32u + 127u) / 128u;       // one warp per chunk, 4 warps per CTA
                const SfqEnc2Chunk *d_e2c = ctx->e2_chunks.as<SfqEnc2Chunk>() + c0;
                uint8_t *abuf = ctx->arena_buf.as<uint8_t>();
                CK(cudaEventRecord(ctx->fork_ev, s));
                CK(cudaStreamWaitEvent(side0, ctx->fork_ev, 0));
                CK(cudaStreamWaitEvent(side1, ctx->fork_ev, 0));
                uint32_t wave_max_nrec = 1;
                for (uint32_t c = c0; c < c0 + nc; c++) wave_max_nrec = std::max(wave_max_nrec, metas[c].nrec);
                if (two_phase && ctx->enc_order == 1) {
                    // the two short head kernels of the quality path get the machine before k_gen_model's long-running
                    // CTAs take its registers (they would otherwise trickle through what is left: 118 + 77 ms against 14 + 10)
                    CK(cudaEventRecord(ctx->wave_ev[WEV * w + 6], sq));
                    TRACED("k_qlt_keys", sq, (k_qlt_keys<<<dim3((wave_max_nrec + SFQ_QK_RECS - 1) / SFQ_QK_RECS, nc), 128, 0, sq>>>(d_text, d_ls, d_metas + c0, ctx->rec_qoff.as<uint32_t>(), e2, d_e2c, level, nc))); LAUNCHED();
                    TRACED("k_qlt_scan", sq, (k_qlt_scan<<<nc, 256, 0, sq>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED();
                    CK(cudaEventRecord(ctx->head_ev, sq));
                    CK(cudaStreamWaitEvent(sg, ctx->head_ev, 0));
                }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 4], sg));
                if (!(ctx->plane_mask & 1)) {}
                else if (two_phase && !gm_table) {
                    const bool gp_smem = (1u << gp_bits) <= SFQ_GP_SMEM_MAX;
                    TRACED("k_gen_keys", sg, (k_gen_keys<<<dim3((wave_max_nrec + SFQ_GC_RECS - 1) / SFQ_GC_RECS, nc), 128, 0, sg>>>(d_text, d_ls, d_metas + c0, ctx->rec_boff.as<uint32_t>(), e2, d_e2c, level, nc))); LAUNCHED();
                    TRACED("k_gen_part", sg, (k_gen_part<<<nc, 32, gp_smem ? (size_t)36 << gp_bits : 0, sg>>>(d_metas + c0, d_arenas + c0, abuf, ws, e2, d_e2c, nc))); LAUNCHED();
                    TRACED("k_gen_replay", sg, (k_gen_replay<<<ctx->sm_count, 32 * SFQ_GR_WARPS, SFQ_GR_SMEM, sg>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED();
                    CK(cudaEventRecord(ctx->gbins_ev, sg));
                    TRACED("k_rc_encode<0>", sg, (k_rc_encode<0><<<((nc + ctx->rc_lanes - 1) / ctx->rc_lanes + ctx->rc_warps - 1) / ctx->rc_warps, 32 * ctx->rc_warps, 0, sg>>>(d_metas + c0, d_arenas + c0, abuf, e2, d_e2c, nc, ctx->rc_lanes))); LAUNCHED();
                } else if (two_phase) {
                    { TraceScope ts_(ctx, "k_gen_model", sg);
                    switch (ctx->gm_variant) {
                    case 1: k_gen_model<4, 8><<<nwarp_blocks, 128, 0, sg>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, e2, d_e2c, level, nc); break;
                    case 2: k_gen_model<8, 8><<<nwarp_blocks, 128, 0, sg>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, e2, d_e2c, level, nc); break;
                    case 3: k_gen_model<8, 1><<<nwarp_blocks, 128, 0, sg>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, e2, d_e2c, level, nc); break;
                    default: k_gen_model<8, 6><<<nwarp_blocks, 128, 0, sg>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, e2, d_e2c, level, nc); break;
                    } }
                    LAUNCHED();
                    TRACED("k_rc_encode<0>", sg, (k_rc_encode<0><<<((nc + ctx->rc_lanes - 1) / ctx->rc_lanes + ctx->rc_warps - 1) / ctx->rc_warps, 32 * ctx->rc_warps, 0, sg>>>(d_metas + c0, d_arenas + c0, abuf, e2, d_e2c, nc, ctx->rc_lanes))); LAUNCHED();
                } else {
                    k_encode<0><<<nb, 32, 0, sg>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, level, nc, lanes); LAUNCHED();
                }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 5], sg));
                if (!(two_phase && ctx->enc_order == 1)) CK(cudaEventRecord(ctx->wave_ev[WEV * w + 6], sq));
                if (!(ctx->plane_mask & 2)) {}
                else if (two_phase) {
                    cudaStream_t q = sq;
                    if (ctx->enc_order != 1) {
                        TRACED("k_qlt_keys", q, (k_qlt_keys<<<dim3((wave_max_nrec + SFQ_QK_RECS - 1) / SFQ_QK_RECS, nc), 128, 0, q>>>(d_text, d_ls, d_metas + c0, ctx->rec_qoff.as<uint32_t>(), e2, d_e2c, level, nc))); LAUNCHED();
                        TRACED("k_qlt_scan", q, (k_qlt_scan<<<nc, 256, 0, q>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED();
                    }
                    if (alias && (ctx->plane_mask & 1)) CK(cudaStreamWaitEvent(q, ctx->gbins_ev, 0));
                    if (ctx->q_scatter1) { TRACED("k_qlt_scatter", q, (k_qlt_scatter<<<nwarp_blocks, 128, 0, q>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED(); }
                    else {
                        TRACED("k_qlt_part1", q, (k_qlt_part1<<<nc, 32, 0, q>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED();
                        TRACED("k_qlt_part2", q, (k_qlt_part2<<<ctx->sm_count * 6, 256, 0, q>>>(d_metas + c0, e2, d_e2c, nc))); LAUNCHED();
                    }
                    CK(cudaEventRecord(ctx->qp2_ev, q));
                    TRACED("k_qlt_model", q, (k_qlt_model<<<ctx->sm_count * 4, SFQ_QM_THREADS, 0, q>>>(d_metas + c0, ws_at(ws, c0), e2, d_e2c, c0))); LAUNCHED();
                    CK(cudaEventRecord(ctx->qtp_ev, q));
                    k_qlt_mark_escapes<<<nc, 256, 0, q>>>(d_metas + c0, e2, d_e2c, nc); LAUNCHED();
                    if (ctx->rc_q4) { TRACED("k_rc_encode_q4", q, (k_rc_encode_q4<<<(nc + 7) / 8, 32, 0, q>>>(d_metas + c0, d_arenas + c0, abuf, e2, d_e2c, nc))); LAUNCHED(); }
                    else { TRACED("k_rc_encode<1>", q, (k_rc_encode<1><<<(nc + ctx->rc_lanes - 1) / ctx->rc_lanes, 32, 0, q>>>(d_metas + c0, d_arenas + c0, abuf, e2, d_e2c, nc, ctx->rc_lanes))); LAUNCHED(); }
                } else {
                    k_encode<1><<<(nc * SFQ_QG + 31) / 32, 32, 0, sq>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, ws, level, nc, lanes); LAUNCHED();
                }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 7], sq));
                if (two_phase && !gm_table && ctx->plane_mask == 7) {
                    // The header encoder is one long chain per chunk whose CTAs hold ~2 KB of shared memory per chunk for the whole wave;
                    // beside it the short throughput kernels of the other two paths run 3-5x slower than alone.  SFQ_ENC_SCHED holds it back:
                    //   1  until the quality models are replayed     2  until both paths' partition passes are done
                    //   4  until the base models are replayed
                    const int sch = ctx->enc_sched;
                    if (sch == 1) CK(cudaStreamWaitEvent(side1, ctx->qtp_ev, 0));
                    if (sch == 2) { CK(cudaStreamWaitEvent(side1, ctx->gbins_ev, 0)); CK(cudaStreamWaitEvent(side1, ctx->qp2_ev, 0)); }
                    if (sch == 4) CK(cudaStreamWaitEvent(side1, ctx->gbins_ev, 0));
                }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 8], side1));
                if (ctx->plane_mask & 4) {
                    const uint32_t rl = ctx->lanes ? ctx->lanes : ctx->enc_rec_lanes ? ctx->enc_rec_lanes : std::min(lanes, 8u);
                    SfqWorkspace wr = ws;
                    if (ctx->rec_global) {                                   // SFQ_REC_GLOBAL=1: the header coder's scratch in global memory (A/B: no shared memory held for the whole wave)
                        CK(ctx->rec_scr.ensure((size_t)nc * sizeof(SfqRecScratch)));
                        wr.rec_scratch = ctx->rec_scr.as<uint8_t>();
                    }
                    TRACED("k_encode<2>", side1, (k_encode<2><<<(nc + rl - 1) / rl, 32, ctx->rec_global ? 0 : rl * sizeof(SfqRecScratch), side1>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, abuf, wr, level, nc, rl))); LAUNCHED();
                }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 9], side1));
                CK(cudaEventRecord(ctx->join_ev[0], side0));
                CK(cudaEventRecord(ctx->join_ev[1], side1));
                CK(cudaStreamWaitEvent(s, ctx->join_ev[0], 0));
                CK(cudaStreamWaitEvent(s, ctx->join_ev[1], 0));
            }
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 2], s));
            k_blob_offsets<<<1, 1024, 0, s>>>(d_metas + c0, d_arenas + c0, d_ls, nc, d_blob_off + c0, d_scal + 1); LAUNCHED();
            k_pack<<<nc, 256, 0, s>>>(d_text, d_ls, d_metas + c0, d_arenas + c0, ctx->arena_buf.as<uint8_t>(), d_blob_off + c0, level,
                                      d_out, out_cap, reinterpret_cast<uint32_t *>(d_scal + 2)); LAUNCHED();
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 3], s));
            host_mark(ctx, "compress: wave enqueued");
        }
        if (part && part->on_enqueued && grow == 0) { const int rc_ = part->on_enqueued(); if (rc_) return rc_; }
        host_mark(ctx, "compress: waves enqueued");
        CK(cudaMemcpyAsync(metas.data(), d_metas, nchunks * sizeof(SfqChunkMeta), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(arenas.data(), d_arenas, nchunks * sizeof(SfqArena), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(blob_off.data(), d_blob_off, nchunks * 8ull, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_small, d_scal, 24, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        host_mark(ctx, "compress: waves done");
        CK(cudaGetLastError());
        trace_dump(ctx, "compress");
        bool again = false;
        for (uint32_t c = 0; c < nchunks; c++) {
            if (metas[c].status == SFQ_E_CAP || metas[c].status == SFQ_E_TABLE) again = true;
            else if (metas[c].status) return status_to_error(ctx, metas[c], c, r0[c]);
        }
        st.ms_clear = st.ms_code = st.ms_pack = st.ms_gen = st.ms_qlt = st.ms_rec = 0;
        for (uint32_t w = 0; w < nwaves; w++) {
            st.ms_clear += ev_ms(ctx->wave_ev[WEV * w], ctx->wave_ev[WEV * w + 1]);
            st.ms_code += ev_ms(ctx->wave_ev[WEV * w + 1], ctx->wave_ev[WEV * w + 2]);
            st.ms_pack += ev_ms(ctx->wave_ev[WEV * w + 2], ctx->wave_ev[WEV * w + 3]);
            st.ms_gen += ev_ms(ctx->wave_ev[WEV * w + 4], ctx->wave_ev[WEV * w + 5]);
            st.ms_qlt += ev_ms(ctx->wave_ev[WEV * w + 6], ctx->wave_ev[WEV * w + 7]);
            st.ms_rec += ev_ms(ctx->wave_ev[WEV * w + 8], ctx->wave_ev[WEV * w + 9]);
        }
        if (!again) { end_cursor = h_small[1]; if (*reinterpret_cast<uint32_t *>(&h_small[2])) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small (need more than %llu bytes)", (unsigned long long)out_cap); break; }
        if (grow >= 6) return fail(ctx, SFQ_ERR_CUDA, "stream arena still too small after 6 doublings");
        st.retries++;
        for (uint32_t c = 0; c < nchunks; c++) { metas[c].status = SFQ_OK; metas[c].g_used = 0; metas[c].q_used = 0; }
        CK(cudaMemcpyAsync(d_metas, metas.data(), nchunks * sizeof(SfqChunkMeta), cudaMemcpyHostToDevice, s));
    }

    // ---- file header + index ("host-side exchange of compressed-size offsets")
    const uint64_t index_off = end_cursor;
    if (part) {
        part->cursor = end_cursor;
        part->index->insert(part->index->end(), blob_off.begin(), blob_off.end());
        part->out_total += out_total;
        CK(cudaEventRecord(ctx->ev[EV_CODE_END], s));
        CK(cudaStreamSynchronize(s));
        *out_n = end_cursor;
    } else {
    if (index_off + nchunks * 8ull > out_cap) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small (need %llu bytes)", (unsigned long long)(index_off + nchunks * 8ull));
    SfqFileHeader fh;
    sfq_file_header_init(&fh, level, n, nchunks, chunk_bytes, index_off, out_total);
    CK(cudaMemcpyAsync(d_out, &fh, sizeof fh, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_out + index_off, blob_off.data(), nchunks * 8ull, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev[EV_CODE_END], s));
    CK(cudaStreamSynchronize(s));
    *out_n = index_off + nchunks * 8ull;
    }
    st.stream_bytes = 0; st.gen_stream_bytes = 0; st.qlt_stream_bytes = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        for (int k = 0; k < SFQ_NSTREAMS; k++) st.stream_bytes += arenas[c].size[k];
        st.gen_stream_bytes += arenas[c].size[SFQ_S_GEN]; st.qlt_stream_bytes += arenas[c].size[SFQ_S_QLT];
    }
    st.in_bytes = n; st.out_bytes = *out_n;
    st.ms_scan = ev_ms(ctx->ev[EV_H2D], ctx->ev[EV_SCAN]);
    st.ms_plan = ev_ms(ctx->ev[EV_SCAN], ctx->ev[EV_PLAN]);
    return 0;
}

// ------------------------------------------------------------------------------------------ decompress
// `h_hdr`/`h_index`/`h_blobs` are host copies of the file header, the blob-offset index and the blob headers.
int decompress_on_device(sfq_ctx *ctx, const uint8_t *d_in, size_t n, const SfqFileHeader &fh,
                         const std::vector<uint64_t> &index, const std::vector<SfqBlobHeader> &blobs,
                         uint8_t *d_out, size_t out_cap, size_t *out_n) {
    cudaStream_t s = ctx->stream;
    cudaStream_t side0 = ctx->serial_roles ? s : ctx->side[0], side1 = ctx->serial_roles ? s : ctx->side[1];
    sfq_stats &st = ctx->st;
    const uint32_t nchunks = (uint32_t)fh.nchunks;
    std::vector<SfqChunkMeta> metas(nchunks);
    std::vector<SfqDecChunk> dcs(nchunks);
    uint64_t nrec = 0, nb = 0, nq = 0, nh = 0, no = 0, max_bases = 0;
    int max_level = 1;
    st.stream_bytes = 0; st.gen_stream_bytes = 0; st.qlt_stream_bytes = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        const SfqBlobHeader &b = blobs[c];
        const uint64_t off = index[c];
        const int bad = sfq_blob_check(&b, off, n);
        if (bad == 2) return fail(ctx, SFQ_ERR_FORMAT, "chunk %u: bad blob header", c);
        SfqChunkMeta &m = metas[c];
        memset(&m, 0, sizeof m);
        m.text_len = b.text_len; m.out_len = b.out_len; m.nrec = b.nrec; m.nbases = b.nbases; m.nquals = b.nquals;
        m.hdr_bytes = b.hdr_bytes; m.llen = b.llen; m.solid = b.solid; m.two_id = b.two_id; m.n_byte = b.n_byte; m.pad = b.pad;
        m.nbig = b.nbig; m.big_bases = b.big_bases; m.big_quals = b.big_quals; m.big_hdr = b.big_hdr;
        SfqDecChunk &d = dcs[c];
        uint64_t o = off + sizeof(SfqBlobHeader);
        d.rec_first_off = o; d.rec_first_len = b.rec_first_len; o += b.rec_first_len;
        for (int k = 0; k < SFQ_NSTREAMS; k++) { d.soff[k] = o; d.ssize[k] = b.ssize[k]; o += b.ssize[k]; st.stream_bytes += b.ssize[k]; }
        st.gen_stream_bytes += b.ssize[SFQ_S_GEN]; st.qlt_stream_bytes += b.ssize[SFQ_S_QLT];
        d.level = (int32_t)b.level;
        d.rec_base = nrec; d.base_plane = nb; d.qual_plane = nq; d.hdr_plane = nh;
        d.base_cap = (uint64_t)b.nbases + b.big_bases; d.qual_cap = (uint64_t)b.nquals + b.big_quals; d.hdr_cap = SFQ_HDR_PLANE(&m);
        nrec += b.nrec; nb += d.base_cap; nq += d.qual_cap; nh += d.hdr_cap; no += b.out_len;
        max_bases = std::max<uint64_t>(max_bases, b.nbases);
        max_level = std::max(max_level, (int)b.level);
        if (bad) return fail(ctx, SFQ_ERR_FORMAT, "chunk %u: inconsistent blob header", c);
    }
    if (no > out_cap) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small (need %llu bytes)", (unsigned long long)no);
    CK(ctx->blob_off.ensure(nchunks * 8ull));          // reused as per-chunk output sizes / offsets
    CK(ctx->scalars.ensure(256));
    CK(ctx->h_small.ensure(4096));
    st.nchunks = nchunks; st.nrecords = nrec; st.nbases = nb; st.nquals = nq;

    CK(ctx->metas.ensure(nchunks * sizeof(SfqChunkMeta)));
    CK(ctx->dchunks.ensure(nchunks * sizeof(SfqDecChunk)));
    CK(ctx->rec_chunk.ensure(nrec * 4 + 64));
    SfqChunkMeta *d_metas = ctx->metas.as<SfqChunkMeta>();
    SfqDecChunk *d_dcs = ctx->dchunks.as<SfqDecChunk>();
    CK(cudaMemcpyAsync(d_metas, metas.data(), nchunks * sizeof(SfqChunkMeta), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_dcs, dcs.data(), nchunks * sizeof(SfqDecChunk), cudaMemcpyHostToDevice, s));
    k_fill_rec_chunk<<<nchunks, 128, 0, s>>>(d_dcs, d_metas, ctx->rec_chunk.as<uint32_t>()); LAUNCHED();      // record -> chunk, for the assemble kernel
    // planes and per-record tables: carved from the scratch arena together with the quality tables (below)
    const size_t fixed_sz[] = {nb + 16, nq + 16, nh + 16, nrec * 4, nrec * 4, nrec * 4, nrec, nrec, nrec * 8, nrec * 8, nrec * 8, nrec * 8};
    DevBuf *fixed_buf[] = {&ctx->bases, &ctx->quals, &ctx->hdrs, &ctx->t_llen, &ctx->t_qlen, &ctx->t_hlen, &ctx->t_pfg, &ctx->t_pfq,
                           &ctx->t_boff, &ctx->t_qoff, &ctx->t_hoff, &ctx->t_ooff};
    size_t fixed_total = 0;
    for (size_t v : fixed_sz) fixed_total += (v + 255) & ~(size_t)255;
    SfqRecTables t{};
    CK(cudaEventRecord(ctx->ev[EV_PLAN], s));

    // all chunks of a container share one table geometry, sized from the blobs' context counts; a table
    // that still fills up (hint missing or wrong) reruns the decode one size larger
    uint64_t max_q = 0, max_g = 0;
    bool have_hints = true;
    for (uint32_t c = 0; c < nchunks; c++) {
        max_q = std::max<uint64_t>(max_q, blobs[c].q_used); max_g = std::max<uint64_t>(max_g, blobs[c].g_used);
        if (!blobs[c].g_used || (!blobs[c].q_used && blobs[c].level > 1)) have_hints = false;
    }
    uint32_t nwaves = 0;
    cudaEvent_t asm0 = ctx->ev[EV_SCAN];    // reused as "assemble start" on the decode timeline
    uint64_t *h_total = ctx->h_small.as<uint64_t>();
    st.retries = 0;
    for (uint32_t grow = 0;; grow++) {
        const uint32_t hbits = sfq_gen_nbuckets(max_level, have_hints ? max_g : max_bases, grow);     // buckets of the decoder's table
        const uint32_t qent = sfq_q_entries(max_level, have_hints ? max_q : 65536, grow);
        const uint32_t cbits = qent;
        const uint64_t gstride = std::max(sfq_gbuckets_bytes(max_level, hbits), sfq_gbuckets_bytes(1, 1));
        const uint64_t qbytes = (uint64_t)std::max(qent, 4096u) * SFQ_L64_WORDS * 4, pbytes = sfq_pwpool_bytes();
        const uint32_t R = pick_resident(ctx, nchunks, gstride + qbytes + pbytes + 4096, ctx->scratch.cap, fixed_total);
        {
            const size_t need = fixed_total + (size_t)R * (qbytes + gstride + pbytes) + 1024;
            cudaError_t e = ctx->scratch.ensure(need);
            if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); release_workspace(ctx); e = ctx->scratch.ensure(need); }
            CK(e);
            size_t off = 0;
            for (int k = 0; k < 12; k++) fixed_buf[k]->carve(ctx->scratch, fixed_sz[k], &off);
            ctx->qtab.carve(ctx->scratch, (size_t)R * qbytes, &off);
            ctx->gtab.carve(ctx->scratch, (size_t)R * gstride, &off);
            ctx->pw.carve(ctx->scratch, (size_t)R * pbytes, &off);
            t.llen = ctx->t_llen.as<uint32_t>(); t.qlen = ctx->t_qlen.as<uint32_t>(); t.hlen = ctx->t_hlen.as<uint32_t>();
            t.pfg = ctx->t_pfg.as<uint8_t>(); t.pfq = ctx->t_pfq.as<uint8_t>();
            t.boff = ctx->t_boff.as<uint64_t>(); t.qoff = ctx->t_qoff.as<uint64_t>(); t.hoff = ctx->t_hoff.as<uint64_t>(); t.ooff = ctx->t_ooff.as<uint64_t>();
        }
        nwaves = (nchunks + R - 1) / R;
        if (ensure_wave_events(ctx, nwaves)) return SFQ_ERR_CUDA;
        st.waves = nwaves; st.resident_chunks = R; st.workspace_bytes = (uint64_t)R * (gstride + qbytes + pbytes);
        SfqWorkspace ws{};
        ws.gtab = ctx->gtab.as<uint8_t>(); ws.gtab_stride = gstride; ws.hbits = hbits;
        ws.qtab = ctx->qtab.as<uint32_t>(); ws.qtab_words = qbytes / 4; ws.cbits = cbits; ws.pw = ctx->pw.as<uint32_t>();
        ws.gen_ahead2 = ctx->gen_ahead2 < 0 ? 1u : (uint32_t)ctx->gen_ahead2;
        for (uint32_t w = 0; w < nwaves; w++) {
            const uint32_t c0 = w * R, nc = std::min(nchunks, c0 + R) - c0;
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 0], s));
            CK(cudaMemsetAsync(ctx->gtab.p, 0, nc * gstride, s));
            CK(cudaMemsetAsync(ctx->qtab.p, 0, nc * qbytes, s));
            CK(cudaMemsetAsync(ctx->pw.p, 0, nc * pbytes, s));
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 1], s));
            k_decode_usr<<<(nc + 31) / 32, 32, 0, s>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, ctx->bases.as<uint8_t>(), ctx->quals.as<uint8_t>(), ctx->hdrs.as<uint8_t>(), nc); LAUNCHED();
            {
                uint32_t lanes = pick_lanes(ctx, nc);
                if (ctx->dec_lanes) lanes = ctx->dec_lanes;
                else if (ctx->dec_fit && nc >= 4096u && !ctx->lanes) {
                    // At most one 4-warp CTA of a thread-per-chunk decoder per SM: 9 481 chunks at 16 per warp are 149 CTAs on
                    // 148 SMs, so one SM always carries two base-decoder CTAs and the launch lasts as long as that SM needs
                    while (lanes < 32u && ((nc + lanes - 1) / lanes + SFQ_DEC_MAXW - 1) / SFQ_DEC_MAXW > (uint32_t)ctx->sm_count) lanes++;
                }
                const unsigned nb = (nc + lanes - 1) / lanes;
                // Launched next to the other two kernels, the base decoder's CTAs are now and then packed onto few SMs by
                // the block scheduler: that call's base decoder runs 3.7x slower while the other two run faster (about one
                // decode call in eight at 10 GB; profiles/README.md, r1e/r1f).  A shared-memory reservation bounds how many
                // CTAs of a kind fit on an SM:
                //   3 (default for large waves)  only the base decoder reserves (at most two of its CTAs per SM): no outlier
                //                                in 7 steps, 1.5 % slower than an outlier-free default step;
                //   1  all three decoders as ~one fat CTA per SM with reservations: no outliers either, but the reservations
                //      take the L1 carve-out with them (quality decoder 860 -> 1 264 ms, header decoder 486 -> 911 ms);
                //   2  the fat CTAs of 1 without reservations: slower and still packs;   0  nothing.
                const int spread = ctx->spread >= 0 ? ctx->spread : nc >= 4096u ? 3 : 0;
                const unsigned dw = spread ? SFQ_DEC_MAXW : ctx->dec_warps;      // (spread 3: only the base decoder reserves)            // warps per CTA of the thread-per-chunk decoders
                uint8_t *pb = ctx->bases.as<uint8_t>(), *pq = ctx->quals.as<uint8_t>(), *ph = ctx->hdrs.as<uint8_t>();
                CK(cudaEventRecord(ctx->fork_ev, s));
                CK(cudaStreamWaitEvent(side0, ctx->fork_ev, 0));
                CK(cudaStreamWaitEvent(side1, ctx->fork_ev, 0));
                // Launch order and priorities: the quality decoder (the longest chain) goes first, on the high-priority stream; the
                // base decoder's ~150 fat CTAs arrive on a machine already holding 600 quality CTAs and are spread one per SM
                // most of the time (about one call in eight they are packed two to an SM and that call's base decoder runs up to
                // 2x slower; the shared-memory reservation above bounds it there).  The other order (SFQ_DEC_ORDER=1: base decoder
                // first, on the empty machine) was measured: the block scheduler then packs them EVERY time (1 378 ms against
                // 681 ms per call, 10 of 10 steps) - it fills an SM to its limit before it moves on.
                cudaStream_t sgd = ctx->dec_gen_first ? side0 : s, sqd = ctx->dec_gen_first ? s : side0;
                auto launch_qlt = [&]() -> int {
                    CK(cudaEventRecord(ctx->wave_ev[WEV * w + 6], sqd));
                    if (!(ctx->plane_mask & 2)) {}
                    else if (ctx->qdec_octets) {
                        // lanes per chunk: 8 while the chains are latency-bound, 4 (twice the chunks per warp, a longer
                        // link) once a wave is large enough for issue slots to be what its warps compete for
                        const uint32_t lpc = ctx->qlpc ? ctx->qlpc : nc >= 4096u ? 4u : 8u;
                        const unsigned qw = (spread == 1 || spread == 2) ? SFQ_QD_MAXW : 2u;               // warps per CTA
                        const unsigned qsm = spread == 1 ? ctx->spread_smem[1] : 0;
                        if (lpc == 4 && ctx->qspec) k_qlt_decode<4, true><<<(nc + 8 * qw - 1) / (8 * qw), 32 * qw, qsm, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pq, nc);
                        else if (lpc == 4 && ctx->qch) k_qlt_decode<4, false, true><<<(nc + 8 * qw - 1) / (8 * qw), 32 * qw, qsm, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pq, nc);
                        else if (lpc == 4) k_qlt_decode<4, false><<<(nc + 8 * qw - 1) / (8 * qw), 32 * qw, qsm, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pq, nc);
                        else if (ctx->qspec) k_qlt_decode<8, true><<<(nc + 4 * qw - 1) / (4 * qw), 32 * qw, qsm, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pq, nc);
                        else k_qlt_decode<8, false><<<(nc + 4 * qw - 1) / (4 * qw), 32 * qw, qsm, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pq, nc);
                        LAUNCHED();
                    }
                    else { k_decode<1><<<(nc + 2 * ctx->qgpw - 1) / (2 * ctx->qgpw), 64, 0, sqd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pb, pq, ph, nc, ctx->qgpw); LAUNCHED(); }
                    CK(cudaEventRecord(ctx->wave_ev[WEV * w + 7], sqd));
                    return 0;
                };
                auto launch_gen = [&]() -> int {
                    CK(cudaEventRecord(ctx->wave_ev[WEV * w + 4], sgd));
                    // base decoder: thread per chunk, `lanes` chunks per warp.  (SFQ_GDEC=1 runs the warp-converged form,
                    // 32 chunks per warp: correct, but its link waits for the slowest of 32 table reads - 45 % slower, kept for A/B.)
                    // (A/B only, off by default: see k_hold_ns)
                    if (!ctx->dec_gen_first && ctx->dec_hold_us && nc >= 4096u && ctx->plane_mask == 7) { k_hold_ns<<<1, 1, 0, sgd>>>((uint64_t)ctx->dec_hold_us * 1000ull); LAUNCHED(); }
                    if (!(ctx->plane_mask & 1)) {}
                    else if (ctx->gdec32 > 0) k_gen_decode32<<<(nc + 31) / 32, 32, 0, sgd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pb, nc);
                    else k_decode<0><<<(nb + dw - 1) / dw, 32 * dw, (spread == 1 || spread == 3) ? ctx->spread_smem[0] : 0, sgd>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pb, pq, ph, nc, lanes);
                    LAUNCHED();
                    if (ctx->plane_mask & 1) { k_gen_exceptions<<<(nc + 31) / 32, 32, 0, sgd>>>(d_in, d_dcs + c0, d_metas + c0, ws, pb, nc); LAUNCHED(); }
                    CK(cudaEventRecord(ctx->wave_ev[WEV * w + 5], sgd));
                    return 0;
                };
                if (ctx->dec_gen_first) { if (launch_gen() || launch_qlt()) return SFQ_ERR_CUDA; }
                else { if (launch_qlt() || launch_gen()) return SFQ_ERR_CUDA; }
                // SFQ_DEC_SCHED=1: the header decoder (154 ms alone, 480 ms beside the other two) starts when the base decoder is done
                if (ctx->dec_sched == 1 && ctx->plane_mask == 7) CK(cudaStreamWaitEvent(side1, ctx->wave_ev[WEV * w + 5], 0));
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 8], side1));
                if (ctx->plane_mask & 4) { k_decode<2><<<(nb + dw - 1) / dw, 32 * dw, spread == 1 ? ctx->spread_smem[2] : 0, side1>>>(d_in, d_dcs + c0, d_metas + c0, ws, t, pb, pq, ph, nc, lanes); LAUNCHED(); }
                CK(cudaEventRecord(ctx->wave_ev[WEV * w + 9], side1));
                CK(cudaEventRecord(ctx->join_ev[0], side0));
                CK(cudaEventRecord(ctx->join_ev[1], side1));
                CK(cudaStreamWaitEvent(s, ctx->join_ev[0], 0));
                CK(cudaStreamWaitEvent(s, ctx->join_ev[1], 0));
            }
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 2], s));
            CK(cudaEventRecord(ctx->wave_ev[WEV * w + 3], s));
        }
        CK(cudaEventRecord(asm0, s));
        // output layout from the decoded lengths; equals the recorded out_len except where the reference
        // itself prints a header differently from how it read it
        uint64_t *d_cout = ctx->blob_off.as<uint64_t>();
        const int plane = ctx->plane_mask == 7 ? 0 : ctx->plane_mask;           // a per-plane hook prints one plane as lines
        k_out_sizes<<<(nchunks + 31) / 32, 32, 0, s>>>(d_dcs, d_metas, t, d_cout, nchunks, plane); LAUNCHED();
        k_scan_u64<<<1, 1024, 0, s>>>(d_cout, nchunks, ctx->scalars.as<uint64_t>()); LAUNCHED();
        k_out_offsets<<<(nchunks + 31) / 32, 32, 0, s>>>(d_dcs, d_metas, t, d_cout, nchunks, plane); LAUNCHED();
        CK(cudaMemcpyAsync(h_total, ctx->scalars.p, 8, cudaMemcpyDeviceToHost, s));
        std::vector<SfqChunkMeta> got(nchunks);
        CK(cudaMemcpyAsync(got.data(), d_metas, nchunks * sizeof(SfqChunkMeta), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        bool again = false;
        for (uint32_t c = 0; c < nchunks; c++) {
            if (got[c].status == SFQ_E_TABLE) again = true;
            else if (got[c].status) return status_to_error(ctx, got[c], c, dcs[c].rec_base);
        }
        if (!again) break;
        if (grow >= 8) return fail(ctx, SFQ_ERR_FORMAT, "context tables still too small after 8 doublings: corrupt container?");
        st.retries++;
        CK(cudaMemcpyAsync(d_metas, metas.data(), nchunks * sizeof(SfqChunkMeta), cudaMemcpyHostToDevice, s));
    }
    no = *h_total;
    if (no > out_cap) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small (need %llu bytes)", (unsigned long long)no);
    if (ctx->plane_mask != 7) {
        const int plane = ctx->plane_mask;
        k_plane_lines<<<(unsigned)((nrec * 32 + 255) / 256), 256, 0, s>>>(d_metas, t, ctx->rec_chunk.as<uint32_t>(),
            plane == 1 ? ctx->bases.as<uint8_t>() : plane == 2 ? ctx->quals.as<uint8_t>() : ctx->hdrs.as<uint8_t>(), plane, d_out, nrec); LAUNCHED();
    } else
    k_assemble<<<(unsigned)(((nrec + SFQ_ASM_RECS - 1) / SFQ_ASM_RECS * 32 + 255) / 256), 256, 0, s>>>(d_dcs, d_metas, t, ctx->rec_chunk.as<uint32_t>(), ctx->bases.as<uint8_t>(),
                                                                   ctx->quals.as<uint8_t>(), ctx->hdrs.as<uint8_t>(), d_out, nrec); LAUNCHED();
    CK(cudaEventRecord(ctx->ev[EV_CODE_END], s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    st.ms_clear = st.ms_code = st.ms_pack = st.ms_gen = st.ms_qlt = st.ms_rec = 0;
    for (uint32_t w = 0; w < nwaves; w++) {
        st.ms_clear += ev_ms(ctx->wave_ev[WEV * w], ctx->wave_ev[WEV * w + 1]);
        st.ms_code += ev_ms(ctx->wave_ev[WEV * w + 1], ctx->wave_ev[WEV * w + 2]);
        st.ms_pack += ev_ms(ctx->wave_ev[WEV * w + 2], ctx->wave_ev[WEV * w + 3]);
        st.ms_gen += ev_ms(ctx->wave_ev[WEV * w + 4], ctx->wave_ev[WEV * w + 5]);
        st.ms_qlt += ev_ms(ctx->wave_ev[WEV * w + 6], ctx->wave_ev[WEV * w + 7]);
        st.ms_rec += ev_ms(ctx->wave_ev[WEV * w + 8], ctx->wave_ev[WEV * w + 9]);
    }
    st.ms_pack += ev_ms(asm0, ctx->ev[EV_CODE_END]);
    st.ms_scan = 0;
    st.ms_plan = ev_ms(ctx->ev[EV_H2D], ctx->ev[EV_PLAN]);
    st.in_bytes = n; st.out_bytes = no;
    *out_n = no;
    return 0;
}

int parse_host_container(sfq_ctx *ctx, const uint8_t *sfq, size_t n, SfqFileHeader &fh,
                         std::vector<uint64_t> &index, std::vector<SfqBlobHeader> &blobs) {
    if (!sfq_is_chunked_container(sfq, n)) return fail(ctx, SFQ_ERR_FORMAT, "not a b200 chunked .sfq container");
    memcpy(&fh, sfq, sizeof fh);
    if (fh.version > SFQ_INTERNAL_VERSION)     // config.cpp:373-377
        return fail(ctx, SFQ_ERR_FORMAT, "compressed with slimfastq version %u. My version is %d. Please upgrade me before decoing", fh.version, SFQ_INTERNAL_VERSION);
    if (fh.nchunks == 0 || fh.nchunks > 0x7fffffffull || fh.index_off > n || n - fh.index_off < fh.nchunks * 8)
        return fail(ctx, SFQ_ERR_FORMAT, "corrupt container index");
    index.resize(fh.nchunks); blobs.resize(fh.nchunks);
    memcpy(index.data(), sfq + fh.index_off, fh.nchunks * 8);
    for (uint64_t c = 0; c < fh.nchunks; c++) {
        if (index[c] > n || n - index[c] < sizeof(SfqBlobHeader)) return fail(ctx, SFQ_ERR_FORMAT, "corrupt container index");
        memcpy(&blobs[c], sfq + index[c], sizeof(SfqBlobHeader));
    }
    return 0;
}

// sfq_compress of a large host buffer, pipelined: the input is cut on the chunk grid into as many parts as it needs coder
// waves anyway (the workspace of all its chunks does not fit the device at once), and while part p is coded, part p+1
// travels to the device and the blobs of part p-1 travel back.  The parts' blobs laid out back to back with one index are
// the one-call container byte for byte (sfq_set_chunk_phase).  Returns -1 when one part is all there is.
int compress_in_parts(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, size_t cap,
                      const uint8_t **out, size_t *out_n) {
    const uint64_t B = !chunk_bytes ? 1ull << 20 : chunk_bytes < 4096 ? 4096 : chunk_bytes;
    int P = ctx->parts;
    // Fractions of the grid slots at which the parts end.  A part costs a fixed time (its chains are latency-bound however
    // few chunks it holds) plus a time per chunk, so the fewest parts win: as many as the input needs coder waves anyway
    // (W; ~12 bytes of workspace per input byte) plus ONE short head part - 8 % of a wave - whose
    // copy-in is the only one nobody can hide and whose coding covers the copy-in of the first full part.
    std::vector<double> ends;
    if (P <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const double have = (double)free_b + (double)ctx->text.cap + (double)ctx->out.cap + (double)ctx->scratch.cap;
        const double avail = 0.92 * (have - 1.2 * (double)n - (double)cap);      // (next to the two text buffers)
        int W = avail <= 0 ? 8 : (int)((12.0 * (double)n + avail - 1) / avail);
        if (W < 1) W = 1;
        if (W > 15) W = 15;
        if (n < (256u << 20) && P == 0) return -1;                 // (SFQ_PARTS=-1: the automatic split whatever the size - tests)
        const double head = ctx->head_frac / W;
        ends.push_back(head);
        for (int w = 1; w <= W; w++) ends.push_back(head + (1.0 - head) * w / W);
        P = W + 1;
    } else {
        if (P > 16) P = 16;
        for (int p = 1; p <= P; p++) ends.push_back((double)p / P);
    }
    // the caller's own buffer may itself be a part of a file (sfq_set_chunk_phase): its grid lines lie at k*B - phase0
    const uint64_t phase0 = ctx->chunk_phase;
    const uint64_t nslots = sfq_slot_count(n, B, phase0);
    if (P <= 1 || nslots < 2 * (uint64_t)P) return -1;
    struct Part { size_t start, end; uint64_t phase; };
    std::vector<Part> parts;
    {
        size_t prev = 0; uint64_t prev_phase = phase0;
        for (int p = 1; p <= P; p++) {
            size_t cut = n; uint64_t ph = 0;
            if (p < P) {
                const uint64_t k = std::min<uint64_t>(nslots, std::max<uint64_t>(1, (uint64_t)((double)nslots * ends[p - 1] + 0.5)));
                const uint64_t line = sfq_slot_target(k, B, phase0);
                cut = line >= n ? n : sfq_record_start_at_or_after(fastq, n, (size_t)line);
                ph = cut - line;
            }
            if (cut > prev) { parts.push_back({prev, cut, prev_phase}); prev = cut; prev_phase = ph; }
        }
    }
    if (parts.size() < 2) return -1;
    ctx->chunk_phase = 0;                                      // consumed: every part below sets its own
    // two text buffers: even parts in the first, odd parts behind it (each sized for its own largest part)
    size_t max_part[2] = {0, 0};
    for (size_t p = 0; p < parts.size(); p++) max_part[p & 1] = std::max(max_part[p & 1], parts[p].end - parts[p].start);
    const size_t tstride = (max_part[0] + 16 + 255) & ~(size_t)255;
    CK(ensure_big(ctx, ctx->text, tstride + max_part[1] + 16));
    CK(ensure_big(ctx, ctx->out, cap));
    cudaStream_t s = ctx->stream;
    uint8_t *d_out = ctx->out.as<uint8_t>();
    // the container is copied out part by part into the pinned result buffer; if that turns out too small (first call
    // on a context) the pieces are skipped and everything is fetched once its size is known
    if (ctx->h_out.cap < n / 6) CK(ctx->h_out.ensure(n / 4 + (1u << 20)));
    bool piecewise = true;
    sfq_stats acc{};
    PartIo io{sizeof(SfqFileHeader), nullptr, 0, nullptr};
    std::vector<uint64_t> index;
    io.index = &index;
    CK(cudaEventRecord(ctx->ev[EV_START], s));
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->ev[EV_START], 0));
    CK(cudaMemcpyAsync(ctx->text.p, fastq + parts[0].start, parts[0].end - parts[0].start, cudaMemcpyHostToDevice, ctx->copy_in));
    CK(cudaEventRecord(ctx->part_ev[0], ctx->copy_in));
    float ms_tail = 0;
    for (size_t p = 0; p < parts.size(); p++) {
        const size_t np = parts[p].end - parts[p].start;
        uint8_t *d_text = ctx->text.as<uint8_t>() + (p & 1) * tstride;
        io.on_enqueued = nullptr;
        if (p + 1 < parts.size()) {                       // the other text buffer is free: the part that used it has been coded
            io.on_enqueued = [ctx, fastq, &parts, p, tstride]() -> int {
                CK(cudaMemcpyAsync(ctx->text.as<uint8_t>() + ((p + 1) & 1) * tstride, fastq + parts[p + 1].start, parts[p + 1].end - parts[p + 1].start,
                                   cudaMemcpyHostToDevice, ctx->copy_in));
                CK(cudaEventRecord(ctx->part_ev[(p + 1) & 1], ctx->copy_in));
                return 0;
            };
        }
        CK(cudaStreamWaitEvent(s, ctx->part_ev[p & 1], 0));
        CK(cudaEventRecord(ctx->ev[EV_H2D], s));
        host_mark(ctx, "parts: next part");
        if (p == 0) { CK(cudaEventSynchronize(ctx->part_ev[0])); acc.ms_h2d = ev_ms(ctx->ev[EV_START], ctx->part_ev[0]); }   // the copy nobody could hide: the first part's
        ctx->chunk_phase = parts[p].phase;
        const uint64_t from = io.cursor;
        size_t on = 0;
        const int rc = compress_on_device(ctx, d_text, np, level, chunk_bytes, d_out, cap, &on, &io);
        if (rc) { cudaStreamSynchronize(ctx->copy_in); cudaStreamSynchronize(ctx->copy_out); return rc; }    // (the caller's buffer may go away)
        {   // per-call statistics: sums over the parts
            const sfq_stats &st = ctx->st;
            acc.nchunks += st.nchunks; acc.nrecords += st.nrecords; acc.nbases += st.nbases; acc.nquals += st.nquals;
            acc.stream_bytes += st.stream_bytes; acc.gen_stream_bytes += st.gen_stream_bytes; acc.qlt_stream_bytes += st.qlt_stream_bytes;
            acc.waves += st.waves; acc.resident_chunks = std::max(acc.resident_chunks, st.resident_chunks); acc.retries += st.retries;
            acc.workspace_bytes = std::max(acc.workspace_bytes, st.workspace_bytes);
            acc.ms_scan += st.ms_scan; acc.ms_plan += st.ms_plan; acc.ms_clear += st.ms_clear; acc.ms_code += st.ms_code; acc.ms_pack += st.ms_pack;
            acc.ms_gen += st.ms_gen; acc.ms_qlt += st.ms_qlt; acc.ms_rec += st.ms_rec;
            acc.kernel_launches = st.kernel_launches;        // (LAUNCHED() keeps counting through the parts)
        }
        if (piecewise && io.cursor + 8ull * index.size() + 4096 > ctx->h_out.cap) piecewise = false;
        if (piecewise) {
            CK(cudaEventRecord(ctx->part_ev[2], s));
            CK(cudaStreamWaitEvent(ctx->copy_out, ctx->part_ev[2], 0));
            CK(cudaMemcpyAsync(ctx->h_out.as<uint8_t>() + from, d_out + from, io.cursor - from, cudaMemcpyDeviceToHost, ctx->copy_out));
        }
    }
    const uint64_t index_off = io.cursor, total = index_off + 8ull * index.size();
    if (total > cap) return fail(ctx, SFQ_ERR_SPACE, "output buffer too small (need %llu bytes)", (unsigned long long)total);
    SfqFileHeader fh;
    sfq_file_header_init(&fh, level > 4 ? 4 : level < 1 ? 1 : level, n, index.size(), B, index_off, io.out_total);
    if (piecewise) {
        CK(cudaEventRecord(ctx->part_ev[3], ctx->copy_out));
        CK(cudaStreamSynchronize(ctx->copy_out));
        memcpy(ctx->h_out.p, &fh, sizeof fh);
        memcpy(ctx->h_out.as<uint8_t>() + index_off, index.data(), 8ull * index.size());
    } else {
        CK(cudaMemcpyAsync(d_out, &fh, sizeof fh, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_out + index_off, index.data(), 8ull * index.size(), cudaMemcpyHostToDevice, s));
        CK(ctx->h_out.ensure(total + total / 8));
        CK(cudaMemcpyAsync(ctx->h_out.p, d_out, total, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ctx->part_ev[3], s));
        CK(cudaStreamSynchronize(s));
    }
    host_mark(ctx, "parts: all copied out");
    ms_tail = ev_ms(ctx->ev[EV_CODE_END], ctx->part_ev[3]);
    acc.in_bytes = n; acc.out_bytes = total;
    acc.ms_d2h = ms_tail > 0 ? ms_tail : 0;                       // the copy nobody could hide: the last part's blobs
    acc.ms_total = ev_ms(ctx->ev[EV_START], ctx->part_ev[3]);
    ctx->st = acc;
    ctx->st.kernel_launches = acc.kernel_launches;
    *out = ctx->h_out.as<uint8_t>();
    *out_n = total;
    return 0;
}

void host_mark(const sfq_ctx *ctx, const char *what) {
    if (!ctx->trace && !ctx->marks) return;
    static std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[sfq host] %10.3f ms  %s\n", ms, what);
}

void begin_call(sfq_ctx *ctx) {
    cudaSetDevice(ctx->device);
    ctx->err.clear();
    memset(&ctx->st, 0, sizeof ctx->st);
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

const char *sfq_version(void) { return "2.04/6 b200"; }

int sfq_create(sfq_ctx **out, int device) {
    if (!out) return SFQ_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SFQ_ERR_CUDA;   // no CPU fallback
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return SFQ_ERR_CUDA;
    if (device >= ndev) return SFQ_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SFQ_ERR_CUDA;
    sfq_ctx *ctx = new sfq_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (const char *e = getenv("SFQ_LANES")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->lanes = (uint32_t)v; }
    if (const char *e = getenv("SFQ_RC_LANES")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->rc_lanes = (uint32_t)v; }
    if (const char *e = getenv("SFQ_ENC_SERIAL")) ctx->serial_encoder = atoi(e) != 0;
    if (const char *e = getenv("SFQ_SPREAD")) ctx->spread = atoi(e);          // 1: fat CTAs + reservation, 2: fat CTAs only
    if (const char *e = getenv("SFQ_DEC_WARPS")) { int v = atoi(e); if (v >= 1 && v <= SFQ_DEC_MAXW) ctx->dec_warps = (uint32_t)v; }
    if (const char *e = getenv("SFQ_ENC_ORDER")) ctx->enc_order = atoi(e);
    if (const char *e = getenv("SFQ_ENC_REC_LANES")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->enc_rec_lanes = (uint32_t)v; }
    if (const char *e = getenv("SFQ_TRACE")) ctx->trace = atoi(e) != 0;
    if (const char *e = getenv("SFQ_GDEC")) ctx->gdec32 = atoi(e) != 0;
    if (const char *e = getenv("SFQ_QLPC")) { int v = atoi(e); if (v == 4 || v == 8) ctx->qlpc = (uint32_t)v; }
    if (const char *e = getenv("SFQ_GM_VARIANT")) ctx->gm_variant = atoi(e);
    if (const char *e = getenv("SFQ_GM_TABLE")) ctx->gm_table = atoi(e) != 0;
    if (const char *e = getenv("SFQ_QSCATTER")) ctx->q_scatter1 = atoi(e) != 0;
    if (const char *e = getenv("SFQ_ENC_PRIO")) ctx->enc_prio_gen = atoi(e) != 0;
    if (const char *e = getenv("SFQ_QSPEC")) ctx->qspec = atoi(e) != 0;
    if (const char *e = getenv("SFQ_ENC_SCHED")) ctx->enc_sched = atoi(e);
    if (const char *e = getenv("SFQ_DEC_SCHED")) ctx->dec_sched = atoi(e);
    if (const char *e = getenv("SFQ_REC_GLOBAL")) ctx->rec_global = atoi(e) != 0;
    if (const char *e = getenv("SFQ_RC_Q4")) ctx->rc_q4 = atoi(e) != 0;
    if (const char *e = getenv("SFQ_DEC_HOLD_US")) { int v = atoi(e); if (v >= 0 && v <= 100000) ctx->dec_hold_us = (uint32_t)v; }
    if (const char *e = getenv("SFQ_RC_WARPS")) { int v = atoi(e); if (v >= 1 && v <= SFQ_RC_MAXW) ctx->rc_warps = (uint32_t)v; }
    if (const char *e = getenv("SFQ_QCH")) ctx->qch = atoi(e) != 0;
    if (const char *e = getenv("SFQ_MARKS")) ctx->marks = atoi(e) != 0;
    if (const char *e = getenv("SFQ_ALIAS")) ctx->alias_steps = atoi(e) != 0;
    if (const char *e = getenv("SFQ_PARTS")) ctx->parts = atoi(e);
    if (const char *e = getenv("SFQ_DEC_LANES")) { int v = atoi(e); if (v >= 1 && v <= 32) ctx->dec_lanes = (uint32_t)v; }
    if (const char *e = getenv("SFQ_DEC_FIT")) ctx->dec_fit = atoi(e) != 0;
    if (const char *e = getenv("SFQ_HEAD_FRAC")) { const double v = atof(e); if (v > 0.0 && v < 1.0) ctx->head_frac = v; }
    if (const char *e = getenv("SFQ_DEC_ORDER")) ctx->dec_gen_first = atoi(e) != 0;
    if (const char *e = getenv("SFQ_GEN_AHEAD2")) ctx->gen_ahead2 = atoi(e) != 0;
    if (const char *e = getenv("SFQ_QDEC")) ctx->qdec_octets = atoi(e) != 0;
    if (const char *e = getenv("SFQ_SERIAL_ROLES")) ctx->serial_roles = atoi(e) != 0;
    if (const char *e = getenv("SFQ_QGPW")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) ctx->qgpw = (uint32_t)v; }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (auto &e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int k = 0; k < 2; k++)
        if (cudaStreamCreateWithPriority(&ctx->side[k], cudaStreamNonBlocking, k == 0 ? prio_hi : prio_lo) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->join_ev[k], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->head_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->gbins_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->qp2_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&ctx->qtp_ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    for (auto &e : ctx->part_ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    if (cudaFuncSetAttribute(k_gen_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SFQ_GR_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_gen_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(36u * SFQ_GP_SMEM_MAX)) != cudaSuccess ||
        cudaFuncSetAttribute(k_encode<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(32 * sizeof(SfqRecScratch))) != cudaSuccess) { delete ctx; return SFQ_ERR_CUDA; }
    {   // shared memory the decoders' CTAs reserve when they are spread (one of each kind per SM fits, two of a kind barely)
        int smem_sm = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
        const unsigned unit = smem_sm > 0 ? (unsigned)smem_sm / 11u : 20u << 10;      // ~20.7 KB on B200 (228 KB per SM): 4 + 3 + 3 units < one SM
        ctx->spread_smem[0] = 4 * unit - 8192; ctx->spread_smem[1] = 3 * unit; ctx->spread_smem[2] = 3 * unit - 1024;
        if (const char *e = getenv("SFQ_GEN_RESERVE_KB")) { const int kb = atoi(e); if (kb >= 0 && kb <= 200) ctx->spread_smem[0] = (unsigned)kb << 10; }   // 112+: one base-decoder CTA per SM
        bool ok = cudaFuncSetAttribute(k_decode<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[0]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_decode<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[2]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_qlt_decode<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[1]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_qlt_decode<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[1]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_qlt_decode<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[1]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_qlt_decode<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[1]) == cudaSuccess &&
                  cudaFuncSetAttribute(k_qlt_decode<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->spread_smem[1]) == cudaSuccess;
        if (!ok) { cudaGetLastError(); ctx->spread_smem[0] = ctx->spread_smem[1] = ctx->spread_smem[2] = 0; }
        // SFQ_GEN_CARVEOUT=<percent>: preferred shared-memory carve-out of the base decoder.  With a carve-out that holds ONE of its
        // CTAs (75 KB reserved + 8 KB static) but not two, the block scheduler cannot pack two onto an SM, and L1 keeps the rest
        // (unlike a 112 KB reservation).  Needs at most one CTA per SM to exist: SFQ_DEC_FIT=1.
        if (const char *e = getenv("SFQ_GEN_CARVEOUT")) {
            const int pct = atoi(e);
            if (pct >= 0 && pct <= 100 && cudaFuncSetAttribute(k_decode<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess) cudaGetLastError();
        }
    }
    *out = ctx;
    return 0;
}

void sfq_destroy(sfq_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->release_all();
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->wave_ev) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) { if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]); if (ctx->join_ev[k]) cudaEventDestroy(ctx->join_ev[k]); }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->head_ev) cudaEventDestroy(ctx->head_ev);
    if (ctx->gbins_ev) cudaEventDestroy(ctx->gbins_ev);
    if (ctx->qp2_ev) cudaEventDestroy(ctx->qp2_ev);
    if (ctx->qtp_ev) cudaEventDestroy(ctx->qtp_ev);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    for (auto &e : ctx->part_ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *sfq_last_error(const sfq_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context (is a CUDA device present?)"; }
int sfq_device_numa_node(int device) {
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, sizeof id, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *p = id; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    char path[96];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", id);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
int sfq_trim(sfq_ctx *ctx) {
    if (!ctx) return SFQ_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return SFQ_ERR_CUDA;
    ctx->release_all();                     // (pointers returned by earlier calls die with the buffers)
    return 0;
}
int sfq_set_max_resident(sfq_ctx *ctx, uint32_t chunks) { if (!ctx) return SFQ_ERR_ARG; ctx->max_resident = chunks; return 0; }
int sfq_set_chunk_phase(sfq_ctx *ctx, uint64_t phase) { if (!ctx) return SFQ_ERR_ARG; ctx->chunk_phase = phase; return 0; }
int sfq_get_stats(const sfq_ctx *ctx, sfq_stats *st) { if (!ctx || !st) return SFQ_ERR_ARG; *st = ctx->st; return 0; }

void *sfq_host_alloc(size_t bytes) { void *p = nullptr; return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
void sfq_host_free(void *p) { if (p) cudaFreeHost(p); }

size_t sfq_compress_bound(size_t n, uint64_t chunk_bytes) {
    if (!chunk_bytes) chunk_bytes = 1ull << 20;
    const uint64_t slots = n / chunk_bytes + 2;
    return (size_t)(n + n / 4 + slots * (sizeof(SfqBlobHeader) + 512 + 12 * SFQ_NSTREAMS + 8) + sizeof(SfqFileHeader) + 4096);
}

int sfq_compress_device(sfq_ctx *ctx, const void *d_fastq, size_t n, int level, uint64_t chunk_bytes,
                        void *d_out, size_t out_cap, size_t *out_n) {
    try {
        if (!ctx || !d_fastq || !d_out || !out_n) return SFQ_ERR_ARG;
        begin_call(ctx);
        CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_H2D], ctx->stream));
        int rc = compress_on_device(ctx, (const uint8_t *)d_fastq, n, level, chunk_bytes, (uint8_t *)d_out, out_cap, out_n);
        if (rc) return rc;
        ctx->st.ms_total = ev_ms(ctx->ev[EV_START], ctx->ev[EV_CODE_END]);
        return 0;
    } catch (const std::bad_alloc &) { return ctx ? fail(ctx, SFQ_ERR_NOMEM, "out of host memory") : SFQ_ERR_NOMEM; }
}

int sfq_compress(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes,
                 const uint8_t **out, size_t *out_n) {
    try {
        if (!ctx || !fastq || !out || !out_n) return SFQ_ERR_ARG;
        begin_call(ctx);
        if (n == 0) return fail(ctx, SFQ_ERR_FASTQ, "no records were found");
        const size_t cap = sfq_compress_bound(n, chunk_bytes);
        {
            const int rc_parts = compress_in_parts(ctx, fastq, n, level, chunk_bytes, cap, out, out_n);
            if (rc_parts != -1) return rc_parts;             // -1: one part only, the plain path below
        }
        CK(ensure_big(ctx, ctx->text, n + 16));
        CK(ensure_big(ctx, ctx->out, cap));
        CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
        CK(cudaMemcpyAsync(ctx->text.p, fastq, n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_H2D], ctx->stream));
        size_t on = 0;
        int rc = compress_on_device(ctx, ctx->text.as<uint8_t>(), n, level, chunk_bytes, ctx->out.as<uint8_t>(), cap, &on);
        if (rc) return rc;
        CK(ctx->h_out.ensure(on));
        CK(cudaMemcpyAsync(ctx->h_out.p, ctx->out.p, on, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->st.ms_h2d = ev_ms(ctx->ev[EV_START], ctx->ev[EV_H2D]);
        ctx->st.ms_d2h = ev_ms(ctx->ev[EV_CODE_END], ctx->ev[EV_D2H]);
        ctx->st.ms_total = ev_ms(ctx->ev[EV_START], ctx->ev[EV_D2H]);
        *out = ctx->h_out.as<uint8_t>();
        *out_n = on;
        return 0;
    } catch (const std::bad_alloc &) { return ctx ? fail(ctx, SFQ_ERR_NOMEM, "out of host memory") : SFQ_ERR_NOMEM; }
}

int sfq_decompressed_size(const uint8_t *sfq, size_t n, uint64_t *out_n, int *level) {
    if (!sfq || !sfq_is_chunked_container(sfq, n)) return SFQ_ERR_FORMAT;
    SfqFileHeader fh;
    memcpy(&fh, sfq, sizeof fh);
    if (out_n) *out_n = fh.out_size;
    if (level) *level = (int)fh.level;
    return 0;
}

int sfq_decompress(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n) {
    try {
        if (!ctx || !sfq || !out || !out_n) return SFQ_ERR_ARG;
        begin_call(ctx);
        SfqFileHeader fh;
        std::vector<uint64_t> index;
        std::vector<SfqBlobHeader> blobs;
        int rc = parse_host_container(ctx, sfq, n, fh, index, blobs);
        if (rc) return rc;
        uint64_t total = 4096;
        for (auto &b : blobs) total += b.out_len + 8ull * b.nrec;      // slack: see SFQ_HDR_PLANE
        CK(ensure_big(ctx, ctx->text, n + 16));          // container bytes
        CK(ensure_big(ctx, ctx->out, total + 16));
        CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
        CK(cudaMemcpyAsync(ctx->text.p, sfq, n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_H2D], ctx->stream));
        size_t on = 0;
        rc = decompress_on_device(ctx, ctx->text.as<uint8_t>(), n, fh, index, blobs, ctx->out.as<uint8_t>(), total, &on);
        if (rc) return rc;
        CK(ctx->h_out_d.ensure(on + 1));
        CK(cudaMemcpyAsync(ctx->h_out_d.p, ctx->out.p, on, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->st.ms_h2d = ev_ms(ctx->ev[EV_START], ctx->ev[EV_H2D]);
        ctx->st.ms_d2h = ev_ms(ctx->ev[EV_CODE_END], ctx->ev[EV_D2H]);
        ctx->st.ms_total = ev_ms(ctx->ev[EV_START], ctx->ev[EV_D2H]);
        *out = ctx->h_out_d.as<uint8_t>();
        *out_n = on;
        return 0;
    } catch (const std::bad_alloc &) { return ctx ? fail(ctx, SFQ_ERR_NOMEM, "out of host memory") : SFQ_ERR_NOMEM; }
}

// ------------------------------------------------------------------------------------------ per-plane test hooks
static int plane_call(sfq_ctx *ctx, int plane, bool enc, const uint8_t *in, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n) {
    if (!ctx) return SFQ_ERR_ARG;
    ctx->plane_mask = plane;
    const int rc = enc ? sfq_compress(ctx, in, n, level, chunk_bytes, out, out_n) : sfq_decompress(ctx, in, n, out, out_n);
    ctx->plane_mask = 7;
    return rc;
}
int sfq_encode_gen_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 1, true, fastq, n, level, chunk_bytes, out, out_n); }
int sfq_encode_qlt_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 2, true, fastq, n, level, chunk_bytes, out, out_n); }
int sfq_encode_rec_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 4, true, fastq, n, level, chunk_bytes, out, out_n); }
int sfq_decode_gen_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 1, false, sfq, n, 0, 0, out, out_n); }
int sfq_decode_qlt_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 2, false, sfq, n, 0, 0, out, out_n); }
int sfq_decode_rec_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n) { return plane_call(ctx, 4, false, sfq, n, 0, 0, out, out_n); }

int sfq_decompress_device(sfq_ctx *ctx, const void *d_sfq, size_t n, void *d_out, size_t out_cap, size_t *out_n) {
    try {
        if (!ctx || !d_sfq || !d_out || !out_n) return SFQ_ERR_ARG;
        begin_call(ctx);
        cudaStream_t s = ctx->stream;
        if (n < sizeof(SfqFileHeader)) return fail(ctx, SFQ_ERR_FORMAT, "not a b200 chunked .sfq container");
        CK(cudaEventRecord(ctx->ev[EV_START], s));
        CK(cudaEventRecord(ctx->ev[EV_H2D], s));
        SfqFileHeader fh;
        CK(cudaMemcpyAsync(&fh, d_sfq, sizeof fh, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (!sfq_is_chunked_container(reinterpret_cast<const uint8_t *>(&fh), sizeof fh)) return fail(ctx, SFQ_ERR_FORMAT, "not a b200 chunked .sfq container");
        if (fh.nchunks == 0 || fh.nchunks > 0x7fffffffull || fh.index_off > n || n - fh.index_off < fh.nchunks * 8)
            return fail(ctx, SFQ_ERR_FORMAT, "corrupt container index");
        std::vector<uint64_t> index(fh.nchunks);
        std::vector<SfqBlobHeader> blobs(fh.nchunks);
        CK(ctx->bhdrs.ensure(fh.nchunks * sizeof(SfqBlobHeader)));
        const uint64_t *d_index = reinterpret_cast<const uint64_t *>((const uint8_t *)d_sfq + fh.index_off);
        if (fh.index_off & 7) {     // unaligned index: stage it through the blob_off buffer
            CK(ctx->blob_off.ensure(fh.nchunks * 8));
            CK(cudaMemcpyAsync(ctx->blob_off.p, (const uint8_t *)d_sfq + fh.index_off, fh.nchunks * 8, cudaMemcpyDeviceToDevice, s));
            d_index = ctx->blob_off.as<uint64_t>();
        }
        k_gather_blob_headers<<<(unsigned)((fh.nchunks + 127) / 128), 128, 0, s>>>((const uint8_t *)d_sfq, d_index, ctx->bhdrs.as<SfqBlobHeader>(), fh.nchunks, n); LAUNCHED();
        CK(cudaMemcpyAsync(index.data(), d_index, fh.nchunks * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(blobs.data(), ctx->bhdrs.p, fh.nchunks * sizeof(SfqBlobHeader), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (uint64_t c = 0; c < fh.nchunks; c++)
            if (index[c] > n || n - index[c] < sizeof(SfqBlobHeader)) return fail(ctx, SFQ_ERR_FORMAT, "corrupt container index");
        int rc = decompress_on_device(ctx, (const uint8_t *)d_sfq, n, fh, index, blobs, (uint8_t *)d_out, out_cap, out_n);
        if (rc) return rc;
        ctx->st.ms_total = ev_ms(ctx->ev[EV_START], ctx->ev[EV_CODE_END]);
        return 0;
    } catch (const std::bad_alloc &) { return ctx ? fail(ctx, SFQ_ERR_NOMEM, "out of host memory") : SFQ_ERR_NOMEM; }
}


// ------------------------------------------------------------------------------------------ reference file format
int sfq_is_reference_file(const uint8_t *p, size_t n) {
    return p && n >= 2 * SFQ_WORM_PAGE && !memcmp(p, SFQ_STAMP, 16) && !sfq_is_chunked_container(p, n);
}

size_t sfq_export_reference_bound(const uint8_t *sfq, size_t n) {
    (void)sfq;
    // every stream rounds up to whole pages and adds one node page per 2047 data pages
    return n + (size_t)(SFQ_NSTREAMS + 4) * 2 * SFQ_WORM_PAGE + n / SFQ_WORM_NODES + SFQ_WORM_PAGE;
}

int sfq_export_reference(const uint8_t *sfq, size_t n, const char *orig_filename, uint8_t *out, size_t out_cap, size_t *out_n) {
    try {
        if (!sfq || !out || !out_n) return SFQ_ERR_ARG;
        if (!sfq_is_chunked_container(sfq, n)) return SFQ_ERR_FORMAT;
        SfqFileHeader fh;
        memcpy(&fh, sfq, sizeof fh);
        if (fh.nchunks != 1) return SFQ_ERR_UNSUPPORTED;               // one reference file = one chunk
        if (fh.index_off > n || n - fh.index_off < 8) return SFQ_ERR_FORMAT;
        uint64_t off;
        memcpy(&off, sfq + fh.index_off, 8);
        if (off > n || n - off < sizeof(SfqBlobHeader)) return SFQ_ERR_FORMAT;
        SfqBlobHeader b;
        memcpy(&b, sfq + off, sizeof b);
        if (b.magic != SFQ_BLOB_MAGIC || off + sfq_blob_size(&b) > n) return SFQ_ERR_FORMAT;
        const uint8_t *p = sfq + off + sizeof b;
        const uint8_t *rec_first = p;
        p += b.rec_first_len;
        const uint8_t *stream[SFQ_NSTREAMS];
        for (int k = 0; k < SFQ_NSTREAMS; k++) { stream[k] = p; p += b.ssize[k]; }
        std::vector<uint8_t> file;
        std::string err;
        if (!sfq_worm_write(b, rec_first, stream, orig_filename, file, err)) return SFQ_ERR_FORMAT;
        if (file.size() > out_cap) return SFQ_ERR_SPACE;
        memcpy(out, file.data(), file.size());
        *out_n = file.size();
        return 0;
    } catch (const std::bad_alloc &) { return SFQ_ERR_NOMEM; }
}

size_t sfq_import_reference_bound(size_t n) { return n + sizeof(SfqFileHeader) + sizeof(SfqBlobHeader) + 0x200 + 64; }

int sfq_import_reference(const uint8_t *ref, size_t n, uint8_t *out, size_t out_cap, size_t *out_n) {
    try {
        if (!ref || !out || !out_n) return SFQ_ERR_ARG;
        std::map<std::string, std::string> info;
        std::map<std::string, std::vector<uint8_t>> streams;
        std::string err;
        if (!sfq_worm_read(ref, n, info, streams, err)) return SFQ_ERR_FORMAT;
        auto num = [&](const char *k, long long dflt) { auto it = info.find(k); return it == info.end() || it->second.empty() ? dflt : atoll(it->second.c_str()); };
        if (num("version", 0) > SFQ_INTERNAL_VERSION) return SFQ_ERR_FORMAT;         // config.cpp:373-377
        const long long orig = num("orig.size", -1), nrec = num("num_records", 0);
        if (orig <= 0 || orig >= 0xFFFFFFF0ll || nrec <= 0 || nrec > 0x7fffffffll) return SFQ_ERR_UNSUPPORTED;
        SfqBlobHeader b;
        memset(&b, 0, sizeof b);
        b.magic = SFQ_BLOB_MAGIC;
        long long level = num("config.level", 2);                                        // config.cpp:363
        b.level = (uint32_t)(level > 4 ? 4 : level < 1 ? 1 : level);
        b.text_len = (uint64_t)orig;
        b.out_len = (uint64_t)orig + 16ull * (uint64_t)nrec + 4096;                      // an upper bound: see SFQ_BLOB_IMPORTED
        b.nrec = (uint32_t)nrec;
        b.nbases = b.nquals = b.hdr_bytes = (uint32_t)orig;                              // upper bounds
        b.llen = (int32_t)num("llen", 0);
        { auto it = info.find("usr.solid"); b.solid = it != info.end() && !it->second.empty() && it->second[0] != '0'; }   // get_bool, config.cpp:124-127
        b.two_id = num("usr.2id", 0) != 0;
        const long long nb = num("gen.N_byte", 'N');
        b.n_byte = (nb && nb != 'N') ? (uint8_t)nb : 0;
        b.pad = SFQ_BLOB_IMPORTED | (num("version", 0) < 5 ? SFQ_BLOB_PRE5 : 0u);         // recs.cpp:397-398: header stream of the older layout
        b.extra_hi = (uint32_t)num("qlt.extra.hi", 0);
        // oversized records: their count is not recorded either; the planes' upper bounds (orig.size) cover them
        b.nbig = 0; b.big_bases = b.big_quals = b.big_hdr = 0;
        const std::string &first = info["rec.first"];
        if (first.size() > 399) return SFQ_ERR_UNSUPPORTED;
        b.rec_first_len = (uint32_t)first.size();
        uint64_t total = sizeof(SfqFileHeader) + sizeof b + first.size();
        for (int k = 0; k < SFQ_NSTREAMS; k++) {
            auto it = streams.find(kSfqStreamNames[k]);
            const size_t sz = it == streams.end() ? 0 : it->second.size();
            if (sz > 0xFFFFFF00ull) return SFQ_ERR_UNSUPPORTED;
            b.ssize[k] = (uint32_t)sz;
            total += sz;
        }
        if (total + 8 > out_cap) return SFQ_ERR_SPACE;
        SfqFileHeader fh;
        sfq_file_header_init(&fh, (int)b.level, (uint64_t)orig, 1, (uint64_t)orig, total, b.out_len);
        uint8_t *p = out;
        memcpy(p, &fh, sizeof fh); p += sizeof fh;
        const uint64_t blob_off = sizeof fh;
        memcpy(p, &b, sizeof b); p += sizeof b;
        memcpy(p, first.data(), first.size()); p += first.size();
        for (int k = 0; k < SFQ_NSTREAMS; k++)
            if (b.ssize[k]) { memcpy(p, streams[kSfqStreamNames[k]].data(), b.ssize[k]); p += b.ssize[k]; }
        memcpy(p, &blob_off, 8); p += 8;
        *out_n = (size_t)(p - out);
        return 0;
    } catch (const std::bad_alloc &) { return SFQ_ERR_NOMEM; }
}


// ------------------------------------------------------------------------------------------ record boundaries (host)
size_t sfq_record_start_at_or_after(const uint8_t *t, size_t n, size_t pos) {
    if (pos == 0) return 0;
    if (!t || pos > n) return n;
    const uint8_t *p = (const uint8_t *)memchr(t + pos - 1, '\n', n - (pos - 1));
    while (p && (size_t)(p - t) + 1 < n) {
        const size_t s = (size_t)(p - t) + 1;
        if (t[s] == '@') {
            const uint8_t *l2 = (const uint8_t *)memchr(t + s, '\n', n - s);
            const uint8_t *l3 = l2 && (size_t)(l2 - t) + 1 < n ? (const uint8_t *)memchr(l2 + 1, '\n', n - (size_t)(l2 + 1 - t)) : nullptr;
            if (l3 && (size_t)(l3 - t) + 1 < n && l3[1] == '+') return s;
        }
        p = (const uint8_t *)memchr(t + s, '\n', n - s);
    }
    return n;
}

size_t sfq_last_record_start(const uint8_t *t, size_t n) {
    if (!t || n == 0) return 0;
    for (size_t w = 1u << 20;; w *= 4) {            // look in a growing window at the end: records are short
        const size_t from = n > w ? n - w : 0;
        size_t cut = 0, s = sfq_record_start_at_or_after(t, n, from ? from : 1);
        while (s < n) { cut = s; s = sfq_record_start_at_or_after(t, n, s + 1); }
        if (cut || from == 0) return cut;
    }
}


size_t sfq_stream_cut(const uint8_t *t, size_t n, uint64_t global_off, uint64_t chunk_bytes, uint64_t *next_phase) {
    if (!t || !n || !chunk_bytes) return 0;
    // grid lines m*B with a local position in (0, n), the last one first; tried a few lines back only
    // (a line without a verifiable record start after it lies in the truncated tail of the buffer)
    const uint64_t last = (global_off + n - 1) / chunk_bytes;
    for (uint64_t m = last, tries = 0; m * chunk_bytes > global_off && tries < 64; m--, tries++) {
        const size_t local = (size_t)(m * chunk_bytes - global_off);
        const size_t s = sfq_record_start_at_or_after(t, n, local);
        if (s < n) {
            if (next_phase) *next_phase = global_off + s - m * chunk_bytes;
            return s;
        }
    }
    return 0;
}

}  // extern "C"
