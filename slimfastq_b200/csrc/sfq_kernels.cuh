// __global__ kernels for sm_100a.  Included once, by sfq_abi.cu.
//
//   scan   k_count_newlines / k_scan_tiles / k_fill_lines   HBM-bound: 16-byte vector loads,
//          byte-compare SIMD (__vcmpeq4), warp-shuffle + shared-memory prefix sums
//   plan   k_chunk_bounds / k_chunk_plan                    binary search + per-chunk framing facts
//   code   k_encode / k_decode (role per blockIdx.y)        one adaptive-coder chain per thread
//   pack   k_blob_sizes / k_pack                            prefix-sum of stream sizes, gather to container
//   build  k_decode_usr / k_out_offsets / k_assemble        planes -> FASTQ text
#pragma once
#include <cuda_runtime.h>
#include "sfq_streams.cuh"
#include "sfq_qlt_group.cuh"
#include "sfq_encode2.cuh"
#include "sfq_plan.cuh"
#include "sfq_container.h"

#define SFQ_SCAN_THREADS 256
#define SFQ_SCAN_ITERS   4
#define SFQ_SCAN_TILE    (SFQ_SCAN_THREADS * SFQ_SCAN_ITERS * 16)     // 16 KiB of text per CTA

// 16 bytes of text starting at `idx` (vectorised when wholly inside the buffer; the buffer base is
// 16-byte aligned, checked by the host).  Bytes past n read as 0.
__device__ __forceinline__ uint4 sfq_load16(const uint8_t *text, uint64_t idx, uint64_t n) {
    if (idx + 16 <= n) return __ldg(reinterpret_cast<const uint4 *>(text + idx));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int b = 0; b < 16; b++)
        if (idx + b < n) w[b >> 2] |= (uint32_t)text[idx + b] << (8 * (b & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
// per-byte 0x80 flags of the newline positions in a 32-bit word
__device__ __forceinline__ uint32_t sfq_nl_flags(uint32_t w) { return __vcmpeq4(w, 0x0a0a0a0au) & 0x80808080u; }
__device__ __forceinline__ uint32_t sfq_nl_count16(uint4 v) {
    return __popc(sfq_nl_flags(v.x)) + __popc(sfq_nl_flags(v.y)) + __popc(sfq_nl_flags(v.z)) + __popc(sfq_nl_flags(v.w));
}

// Pass 1: newlines per 16 KiB tile.
__global__ void __launch_bounds__(SFQ_SCAN_THREADS)
k_count_newlines(const uint8_t *__restrict__ text, uint64_t n, uint32_t *__restrict__ tile_counts) {
    const uint64_t base = (uint64_t)blockIdx.x * SFQ_SCAN_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < SFQ_SCAN_ITERS; it++) {
        const uint64_t idx = base + ((uint64_t)it * SFQ_SCAN_THREADS + threadIdx.x) * 16;
        if (idx < n) c += sfq_nl_count16(sfq_load16(text, idx, n));
    }
    __shared__ uint32_t wsum[SFQ_SCAN_THREADS / 32];
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SFQ_SCAN_THREADS / 32; w++) t += wsum[w];
        tile_counts[blockIdx.x] = t;
    }
}

// Exclusive prefix sum of the tile counts (single CTA; ntiles is ~6e5 for 10 GB).
__global__ void __launch_bounds__(1024)
k_scan_tiles(const uint32_t *__restrict__ counts, uint64_t *__restrict__ prefix, uint64_t ntiles, uint64_t *total) {
    __shared__ uint64_t part[1024];
    const uint64_t per = (ntiles + 1023) / 1024;
    const uint64_t lo = (uint64_t)threadIdx.x * per, hi = lo + per < ntiles ? lo + per : ntiles;
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; i++) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 1024; i++) { uint64_t t = part[i]; part[i] = run; run += t; }
        *total = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint64_t i = lo; i < hi; i++) { prefix[i] = run; run += counts[i]; }
}

// Pass 2: line_start[k+1] = offset after the k-th newline; line_start[0] = 0.
__global__ void __launch_bounds__(SFQ_SCAN_THREADS)
k_fill_lines(const uint8_t *__restrict__ text, uint64_t n, const uint64_t *__restrict__ tile_prefix,
             uint64_t *__restrict__ line_start) {
    const uint64_t base = (uint64_t)blockIdx.x * SFQ_SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = SFQ_SCAN_THREADS / 32;
    __shared__ uint32_t wtot[SFQ_SCAN_ITERS][NW];
    uint4 v[SFQ_SCAN_ITERS];
    uint32_t incl[SFQ_SCAN_ITERS], cnt[SFQ_SCAN_ITERS];
#pragma unroll
    for (int it = 0; it < SFQ_SCAN_ITERS; it++) {
        const uint64_t idx = base + ((uint64_t)it * SFQ_SCAN_THREADS + threadIdx.x) * 16;
        v[it] = idx < n ? sfq_load16(text, idx, n) : make_uint4(0, 0, 0, 0);
        uint32_t c = sfq_nl_count16(v[it]);
        cnt[it] = c;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += t; }
        incl[it] = c;
        if (lane == 31) wtot[it][warp] = c;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    uint64_t run = tile_prefix[blockIdx.x];
#pragma unroll
    for (int it = 0; it < SFQ_SCAN_ITERS; it++) {
        uint32_t before = 0, all = 0;
        for (int w = 0; w < NW; w++) { uint32_t t = wtot[it][w]; if (w < warp) before += t; all += t; }
        uint64_t rank = run + before + (incl[it] - cnt[it]);
        const uint64_t idx = base + ((uint64_t)it * SFQ_SCAN_THREADS + threadIdx.x) * 16;
        const uint32_t w4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t m = sfq_nl_flags(w4[k]);
            while (m) {
                const int b = (__ffs(m) - 1) >> 3;
                line_start[++rank] = idx + 4 * k + b + 1;
                m &= m - 1;
            }
        }
        run += all;
    }
}

// rec_begin[c] = first record starting at or after grid slot c (c in [0, nslots]); see sfq_slot_target.
__global__ void k_chunk_bounds(const uint64_t *__restrict__ ls, uint64_t nrec_total, uint64_t chunk_bytes,
                               uint64_t nslots, uint64_t phase, uint64_t *__restrict__ rec_begin) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nslots) return;
    rec_begin[c] = c == nslots ? nrec_total : sfq_first_record_at(ls, nrec_total, sfq_slot_target(c, chunk_bytes, phase));
}
__global__ void k_chunk_plan(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls,
                             const uint64_t *__restrict__ r0, const uint64_t *__restrict__ r1,
                             SfqChunkMeta *metas, uint32_t nchunks, uint32_t *rec_qoff, uint32_t *rec_boff) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    sfq_plan_chunk(text, ls, r0[c], r1[c], &metas[c], rec_qoff + r0[c], rec_boff + r0[c]);
}

// One thread = one chunk-stream.  ROLE 0 = gen (+gen.Ns/Nn), 1 = qlt, 2 = rec (+rec.x, usr.*); the three
// roles of a wave are separate kernels launched on three streams so they overlap on the device and
// each gets its own register allocation.
// `lanes` = chunk-streams per warp (<= 32): fewer lanes per warp means more warps, so one lane's
// cache miss or slow path stalls fewer neighbours.
template <int ROLE>
__global__ void __launch_bounds__(32)
k_encode(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, SfqChunkMeta *metas,
         SfqArena *arenas, uint8_t *arena_buf, SfqWorkspace ws, int level, uint32_t nchunks, uint32_t lanes) {
    if (ROLE == 1) {                       // lane-cooperative: SFQ_QG lanes per chunk
        const uint32_t c = blockIdx.x * (32 / SFQ_QG) + threadIdx.x / SFQ_QG;
        if (c >= nchunks || metas[c].status != SFQ_OK) return;
        SfqQGroup g;
        g.lane = threadIdx.x % SFQ_QG; g.gbase = threadIdx.x & ~(SFQ_QG - 1u); g.gmask = ((1u << SFQ_QG) - 1u) << g.gbase;
        sfq_qlt_encode_group(text, ls, &metas[c], level, ws.qtab + (size_t)c * ws.qtab_words, ws.cbits,
                             ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS, arena_buf, &arenas[c], g);
        return;
    }
    const uint32_t c = blockIdx.x * lanes + threadIdx.x;
    if (threadIdx.x >= lanes || c >= nchunks) return;
    SfqChunkMeta *m = &metas[c];
    if (m->status != SFQ_OK) return;
    uint32_t *pw = ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS;
    if (ROLE == 0) sfq_gen_encode_chunk(text, ls, m, level, ws.gtab + (size_t)c * ws.gtab_stride, ws.hbits, pw, arena_buf, &arenas[c]);
    else {
        extern __shared__ uint64_t rec_smem[];             // one SfqRecScratch per chunk-stream of the warp
        SfqRecScratch *scr = ws.rec_scratch ? reinterpret_cast<SfqRecScratch *>(ws.rec_scratch) + c : reinterpret_cast<SfqRecScratch *>(rec_smem) + threadIdx.x;
        sfq_rec_encode_chunk(text, ls, m, pw, arena_buf, &arenas[c], scr);
    }
}

// `rec.first` (recs.cpp:68-75) = id line of the first record that went through the models; none if every record was oversized
__device__ __forceinline__ uint32_t sfq_rec_first_len(const SfqChunkMeta &m, const uint64_t *ls) {
    if (m.first_coded >= m.nrec) return 0;
    const uint64_t *l = ls + m.line0 + 4ull * m.first_coded;
    return (uint32_t)(l[1] - l[0] - 2);
}
// Blob sizes of a wave and their exclusive prefix from *cursor (single CTA).
__global__ void __launch_bounds__(1024)
k_blob_offsets(const SfqChunkMeta *__restrict__ metas, const SfqArena *__restrict__ arenas,
               const uint64_t *__restrict__ ls, uint32_t nchunks, uint64_t *blob_off, uint64_t *cursor) {
    __shared__ uint64_t part[1024];
    const uint32_t per = (nchunks + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < nchunks ? lo + per : nchunks;
    uint64_t s = 0;
    for (uint32_t c = lo; c < hi; c++) {
        uint64_t b = sizeof(SfqBlobHeader) + sfq_rec_first_len(metas[c], ls);
        for (int k = 0; k < SFQ_NSTREAMS; k++) b += arenas[c].size[k];
        blob_off[c] = b;
        s += b;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = *cursor;
        for (int i = 0; i < 1024; i++) { uint64_t t = part[i]; part[i] = run; run += t; }
        *cursor = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint32_t c = lo; c < hi; c++) { uint64_t b = blob_off[c]; blob_off[c] = run; run += b; }
}

// One CTA per chunk: header + rec.first + streams -> container.  Sets *overflow if out is too small.
__global__ void __launch_bounds__(256)
k_pack(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, const SfqChunkMeta *__restrict__ metas,
       const SfqArena *__restrict__ arenas, const uint8_t *__restrict__ arena_buf,
       const uint64_t *__restrict__ blob_off, int level, uint8_t *out, uint64_t out_cap, uint32_t *overflow) {
    const uint32_t c = blockIdx.x;
    const SfqChunkMeta &m = metas[c];
    const SfqArena &a = arenas[c];
    __shared__ SfqBlobHeader h;
    const uint32_t rfl = sfq_rec_first_len(m, ls);
    if (threadIdx.x == 0) {
        h.magic = SFQ_BLOB_MAGIC; h.level = (uint32_t)level; h.text_len = m.text_len; h.out_len = m.out_len;
        h.nrec = m.nrec; h.nbases = m.nbases; h.nquals = m.nquals; h.hdr_bytes = m.hdr_bytes; h.llen = m.llen;
        h.solid = m.solid; h.two_id = m.two_id; h.n_byte = m.n_byte; h.pad = 0; h.extra_hi = m.extra_hi;
        h.rec_first_len = rfl; h.q_used = m.q_used; h.g_used = m.g_used;
        h.nbig = m.nbig; h.big_bases = m.big_bases; h.big_quals = m.big_quals; h.big_hdr = m.big_hdr;
        for (int k = 0; k < SFQ_NSTREAMS; k++) h.ssize[k] = a.size[k];
    }
    __syncthreads();
    uint64_t total = sizeof(SfqBlobHeader) + rfl;
    for (int k = 0; k < SFQ_NSTREAMS; k++) total += a.size[k];
    uint64_t o = blob_off[c];
    if (o + total > out_cap) { if (threadIdx.x == 0) atomicExch(overflow, 1u); return; }
    const uint8_t *hp = reinterpret_cast<const uint8_t *>(&h);
    for (uint32_t i = threadIdx.x; i < sizeof(SfqBlobHeader); i += blockDim.x) out[o + i] = hp[i];
    o += sizeof(SfqBlobHeader);
    const uint8_t *rf = text + (rfl ? ls[m.line0 + 4ull * m.first_coded] + 1 : 0);
    for (uint32_t i = threadIdx.x; i < rfl; i += blockDim.x) out[o + i] = rf[i];
    o += rfl;
    for (int k = 0; k < SFQ_NSTREAMS; k++) {
        const uint8_t *src = arena_buf + a.off[k];
        const uint32_t sz = a.size[k];
        for (uint32_t i = threadIdx.x; i < sz; i += blockDim.x) out[o + i] = src[i];
        o += sz;
    }
}

// ---------------------------------------------------------------------------- decode side
// What the decoders need to know about a chunk of the wave (built by the host from blob headers).
struct SfqDecChunk {
    uint64_t soff[SFQ_NSTREAMS];    // absolute offsets of the streams in the container buffer
    uint32_t ssize[SFQ_NSTREAMS];
    uint64_t rec_first_off;
    uint32_t rec_first_len;
    int32_t  level;
    uint64_t rec_base;              // first slot of the chunk in the per-record tables
    uint64_t base_plane, qual_plane, hdr_plane;   // plane offsets of the chunk
    uint64_t base_cap, qual_cap, hdr_cap;         // bytes of each plane that belong to the chunk
};
struct SfqRecTables {
    uint32_t *llen, *qlen, *hlen;
    uint8_t *pfg, *pfq;
    uint64_t *boff, *qoff, *hoff, *ooff;
};

#include "sfq_qlt_dec.cuh"
#include "sfq_gen_dec.cuh"

// rec_chunk[k] = chunk of record k (one CTA per chunk)
__global__ void __launch_bounds__(128)
k_fill_rec_chunk(const SfqDecChunk *__restrict__ dc, const SfqChunkMeta *__restrict__ metas, uint32_t *rec_chunk) {
    const uint32_t c = blockIdx.x;
    uint32_t *p = rec_chunk + dc[c].rec_base;
    for (uint32_t r = threadIdx.x; r < metas[c].nrec; r += blockDim.x) p[r] = c;
}

__global__ void __launch_bounds__(32)
k_decode_usr(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, SfqChunkMeta *metas,
             SfqWorkspace ws, SfqRecTables t, uint8_t *bases, uint8_t *quals, uint8_t *hdrs, uint32_t nchunks) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const SfqDecChunk &d = dc[c];
    SfqChunkMeta *m = &metas[c];
    uint32_t *pw = ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS;
    sfq_usr_decode_chunk(in, d.ssize, d.soff, m, pw, t.llen + d.rec_base, t.qlen + d.rec_base,
                         t.pfg + d.rec_base, t.pfq + d.rec_base, t.hlen + d.rec_base, t.hoff + d.rec_base, t.boff + d.rec_base, t.qoff + d.rec_base,
                         hdrs + d.hdr_plane, d.hdr_cap, bases + d.base_plane, d.base_cap, quals + d.qual_plane, d.qual_cap);
    // oversized records' lines lie at the front of the chunk's planes, the coded records follow
    uint64_t b = d.base_plane + m->big_bases, q = d.qual_plane + m->big_quals;
    for (uint32_t r = 0; r < m->nrec; r++) {
        const uint64_t k = d.rec_base + r;
        if (t.llen[k] & SFQ_BIG_BIT) { t.boff[k] += d.base_plane; t.qoff[k] += d.qual_plane; continue; }
        t.boff[k] = b; t.qoff[k] = q;
        b += t.llen[k]; q += t.qlen[k];
    }
    if (m->status == SFQ_OK && (b > d.base_plane + d.base_cap || q > d.qual_plane + d.qual_cap)) m->status = SFQ_E_CORRUPT;
}

// After the base decoder of a wave: apply gen.Ns / gen.Nn to the decoded base planes (one thread per chunk).
__global__ void __launch_bounds__(32)
k_gen_exceptions(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, const SfqChunkMeta *__restrict__ metas,
                 SfqWorkspace ws, uint8_t *bases, uint32_t nchunks) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks || metas[c].status != SFQ_OK) return;
    const SfqDecChunk &d = dc[c];
    if (d.ssize[SFQ_S_GEN_NS] == 0 && d.ssize[SFQ_S_GEN_NN] == 0) return;
    // (positions count coded bases; the coded records' lines follow the oversized records' at the front of the plane)
    sfq_gen_apply_exceptions(in, d.ssize, d.soff, &metas[c], ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS, bases + d.base_plane + metas[c].big_bases);
}

// Holds a stream back for `ns` nanoseconds of device time (one thread).  An experiment on the packed-CTA outlier of the decode
// wave (about one call in thirty the base decoder's CTAs sit two to an SM on half the SMs and that kernel runs 2.1x slower,
// 1 423 against 671 ms): launched alone on an empty machine the block scheduler packs them every time (round 2, r2m), and held back
// until the other two decoders are in place it ALSO packs them every time (r2ah: 60 of 60 calls) - the even spread of the
// default launch comes from the three grids being distributed at the same moment.  Kept for A/B (SFQ_DEC_HOLD_US), off by default.
__global__ void k_hold_ns(uint64_t ns) {
    uint64_t t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

#define SFQ_DEC_MAXW 4                 // warps per CTA of the thread-per-chunk decoders (1..4; more per CTA = fewer, fatter CTAs)
template <int ROLE>
__global__ void __launch_bounds__(ROLE == 1 ? 64 : 32 * SFQ_DEC_MAXW)
k_decode(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, SfqChunkMeta *metas,
         SfqWorkspace ws, SfqRecTables t, uint8_t *bases, uint8_t *quals, uint8_t *hdrs, uint32_t nchunks, uint32_t lanes) {
    __shared__ uint32_t lut[ROLE == 0 ? SFQ_B2_LUT : 1];       // reciprocals of the 4-symbol model's totals
    __shared__ uint4 cells[ROLE == 0 ? 64 * SFQ_DEC_MAXW : 1]; // bucket look-ahead, two 16-byte cells per thread
    if (ROLE == 0) { sfq_b2_lut_fill(lut, threadIdx.x, blockDim.x); __syncthreads(); }
    if (ROLE == 1) {
        // `lanes` = groups (chunks) per warp: fewer groups per warp means fewer groups waiting on each
        // other's divergent branches and memory round trips (the chain is latency-bound, lanes are cheap)
        const uint32_t gw = (threadIdx.x & 31u) / SFQ_QG;
        const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + gw;
        if (gw >= lanes || c >= nchunks || metas[c].status != SFQ_OK) return;
        const SfqDecChunk &d = dc[c];
        SfqQGroup g;
        g.lane = threadIdx.x % SFQ_QG; g.gbase = threadIdx.x & 31u & ~(SFQ_QG - 1u); g.gmask = ((1u << SFQ_QG) - 1u) << g.gbase;
        sfq_qlt_decode_group(in, d.ssize, d.soff, &metas[c], d.level, ws.qtab + (size_t)c * ws.qtab_words, ws.cbits,
                             ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS, t.qlen + d.rec_base, t.qoff + d.rec_base, quals, g);
        return;
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + lane;
    if (lane >= lanes || c >= nchunks) return;
    const SfqDecChunk &d = dc[c];
    SfqChunkMeta *m = &metas[c];
    if (m->status != SFQ_OK) return;
    uint32_t *pw = ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS;
    if (ROLE == 0) {
        SfqStage stage;
        stage.cell0 = &cells[threadIdx.x]; stage.cell1 = &cells[blockDim.x + threadIdx.x];
        sfq_gen_decode_chunk(in, d.ssize, d.soff, m, d.level, ws.gtab + (size_t)c * ws.gtab_stride, ws.hbits, pw,
                             t.llen + d.rec_base, t.boff + d.rec_base, bases, lut, stage, ws.gen_ahead2);
    }
    else {
        sfq_rec_decode_chunk(in, d.ssize, d.soff, m, pw, in + d.rec_first_off, d.rec_first_len,
                             hdrs + d.hdr_plane, SFQ_HDR_PLANE(m), t.hlen + d.rec_base, t.hoff + d.rec_base);
        for (uint32_t r = 0; r < m->nrec; r++) t.hoff[d.rec_base + r] += d.hdr_plane;
    }
}

// Output layout (UsrLoad::save line layout, usrs.cpp:512-529), from the decoded lengths: bytes per
// chunk, exclusive prefix over all chunks of the container, then the offset of every record.
// `plane` 0 = the FASTQ text; 1 / 2 / 4 = one decoded plane as lines (per-plane test hooks): the record's base line,
// quality line or id line (an oversized record's id + '+' line block) followed by a newline
__device__ __forceinline__ uint64_t sfq_rec_out_len(const SfqChunkMeta &m, const SfqRecTables &t, uint64_t k, int plane = 0) {
    if (plane) return 1ull + ((plane == 1 ? t.llen[k] : plane == 2 ? t.qlen[k] : t.hlen[k]) & ~SFQ_BIG_BIT);
    const uint32_t s = m.solid ? 1 : 0;
    // oversized: '@' id '\n' bases '\n' '+'line '\n' quals '\n', where hlen counts id + '\n' + '+'line
    if (t.hlen[k] & SFQ_BIG_BIT) return 4ull + (t.hlen[k] & ~SFQ_BIG_BIT) + (t.llen[k] & ~SFQ_BIG_BIT) + (t.qlen[k] & ~SFQ_BIG_BIT);
    return 1ull + t.hlen[k] + 1 + s + t.llen[k] + 1 + 1 + (m.two_id ? t.hlen[k] : 0) + 1 + s + t.qlen[k] + 1;
}
__global__ void __launch_bounds__(32)
k_out_sizes(const SfqDecChunk *__restrict__ dc, const SfqChunkMeta *__restrict__ metas, SfqRecTables t,
            uint64_t *chunk_out, uint32_t nchunks, int plane) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    uint64_t o = 0;
    if (metas[c].status == SFQ_OK)
        for (uint32_t r = 0; r < metas[c].nrec; r++) o += sfq_rec_out_len(metas[c], t, dc[c].rec_base + r, plane);
    chunk_out[c] = o;
}
__global__ void __launch_bounds__(1024)
k_scan_u64(uint64_t *v, uint32_t n, uint64_t *total) {        // in-place exclusive scan, single CTA
    __shared__ uint64_t part[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    uint64_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += v[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 1024; i++) { uint64_t x = part[i]; part[i] = run; run += x; }
        *total = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { uint64_t x = v[i]; v[i] = run; run += x; }
}
__global__ void __launch_bounds__(32)
k_out_offsets(const SfqDecChunk *__restrict__ dc, const SfqChunkMeta *__restrict__ metas, SfqRecTables t,
              const uint64_t *__restrict__ chunk_out_off, uint32_t nchunks, int plane) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks || metas[c].status != SFQ_OK) return;
    uint64_t o = chunk_out_off[c];
    for (uint32_t r = 0; r < metas[c].nrec; r++) {
        const uint64_t k = dc[c].rec_base + r;
        t.ooff[k] = o;
        o += sfq_rec_out_len(metas[c], t, k, plane);
    }
}

// One warp per record: planes -> FASTQ text, applying the "quality '!' means N" rule of
// GenLoad::normalize_gen (gens.cpp:200-213) with the flags left by sfq_gen_decode_chunk.
// (SFQ_ASM_RECS consecutive records per warp: a warp that lives for one 340-byte record spends its life being launched)
#define SFQ_ASM_RECS 8u
__device__ __forceinline__ void sfq_assemble_record(const SfqChunkMeta *__restrict__ metas, const SfqRecTables &t,
           const uint32_t *__restrict__ rec_chunk, const uint8_t *__restrict__ bases,
           const uint8_t *__restrict__ quals, const uint8_t *__restrict__ hdrs, uint8_t *out, uint64_t k, uint32_t lane) {
    const SfqChunkMeta &m = metas[rec_chunk[k]];
    if (m.status != SFQ_OK) return;
    const uint32_t hlen = t.hlen[k], llen = t.llen[k], qlen = t.qlen[k];
    const uint8_t *h = hdrs + t.hoff[k], *b = bases + t.boff[k], *q = quals + t.qoff[k];
    const uint8_t nbyte = m.n_byte ? m.n_byte : (uint8_t)'N';
    uint8_t *o = out + t.ooff[k];
    if (hlen & SFQ_BIG_BIT) {
        // an oversized record, verbatim (usrs.cpp:473-485): the block holds the id line, '\n', the '+' line
        const uint32_t hb = hlen & ~SFQ_BIG_BIT, bl = llen & ~SFQ_BIG_BIT, ql = qlen & ~SFQ_BIG_BIT;
        uint32_t cut = hb;                                   // position of the '\n' that ends the id line
        for (uint32_t i0 = 0; i0 < hb && cut == hb; i0 += 32) {
            const unsigned nl = __ballot_sync(0xffffffffu, i0 + lane < hb && h[i0 + lane] == '\n');
            if (nl) cut = i0 + (uint32_t)__ffs(nl) - 1;
        }
        if (lane == 0) o[0] = '@';
        for (uint32_t i = lane; i <= cut && i < hb; i += 32) o[1 + i] = h[i];       // id and its newline
        o += 1 + cut + 1;
        for (uint32_t i = lane; i < bl; i += 32) o[i] = b[i];
        if (lane == 0) o[bl] = '\n';
        o += bl + 1;
        const uint32_t pl = hb - cut - 1;
        for (uint32_t i = lane; i < pl; i += 32) o[i] = h[cut + 1 + i];
        if (lane == 0) o[pl] = '\n';
        o += pl + 1;
        for (uint32_t i = lane; i < ql; i += 32) o[i] = q[i];
        if (lane == 0) o[ql] = '\n';
        return;
    }
    if (lane == 0) o[0] = '@';
    for (uint32_t i = lane; i < hlen; i += 32) o[1 + i] = h[i];
    o += 1 + hlen;
    if (lane == 0) o[0] = '\n';
    o += 1;
    if (m.solid) { if (lane == 0) o[0] = t.pfg[k]; o += 1; }
    for (uint32_t i = lane; i < llen; i += 32) {
        uint8_t c = b[i];
        const uint8_t qc = i < qlen ? q[i] : (uint8_t)40;
        if (c & 0x80) c &= 0x7f; else if (qc == '!') c = nbyte;
        o[i] = c;
    }
    o += llen;
    if (lane == 0) { o[0] = '\n'; o[1] = '+'; }
    o += 2;
    if (m.two_id) { for (uint32_t i = lane; i < hlen; i += 32) o[i] = h[i]; o += hlen; }
    if (lane == 0) o[0] = '\n';
    o += 1;
    if (m.solid) { if (lane == 0) o[0] = t.pfq[k]; o += 1; }
    for (uint32_t i = lane; i < qlen; i += 32) o[i] = q[i];
    if (lane == 0) o[qlen] = '\n';
}
__global__ void __launch_bounds__(256)
k_assemble(const SfqDecChunk *__restrict__ dc, const SfqChunkMeta *__restrict__ metas, SfqRecTables t,
           const uint32_t *__restrict__ rec_chunk, const uint8_t *__restrict__ bases,
           const uint8_t *__restrict__ quals, const uint8_t *__restrict__ hdrs, uint8_t *out, uint64_t nrec) {
    (void)dc;
    const uint64_t k0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * SFQ_ASM_RECS;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t k = k0; k < k0 + SFQ_ASM_RECS && k < nrec; k++)
        sfq_assemble_record(metas, t, rec_chunk, bases, quals, hdrs, out, k, lane);
}

// Per-plane test hooks: one warp per record copies the record's line of ONE decoded plane (bases with the exception
// lists applied but without the "quality '!' means N" rule, which needs the other plane; qualities; id lines).
__global__ void __launch_bounds__(256)
k_plane_lines(const SfqChunkMeta *__restrict__ metas, SfqRecTables t, const uint32_t *__restrict__ rec_chunk,
              const uint8_t *__restrict__ plane_buf, int plane, uint8_t *out, uint64_t nrec) {
    const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (k >= nrec || metas[rec_chunk[k]].status != SFQ_OK) return;
    const uint32_t len = (plane == 1 ? t.llen[k] : plane == 2 ? t.qlen[k] : t.hlen[k]) & ~SFQ_BIG_BIT;
    const uint8_t *src = plane_buf + (plane == 1 ? t.boff[k] : plane == 2 ? t.qoff[k] : t.hoff[k]);
    uint8_t *o = out + t.ooff[k];
    for (uint32_t i = lane; i < len; i += 32) o[i] = plane == 1 ? (uint8_t)(src[i] & 0x7fu) : src[i];
    if (lane == 0) o[len] = '\n';
}

// Blob headers of a device-resident container -> contiguous array (for the host to plan decode).
__global__ void k_gather_blob_headers(const uint8_t *__restrict__ in, const uint64_t *__restrict__ blob_off,
                                      SfqBlobHeader *hdrs, uint64_t nchunks, uint64_t in_size) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    uint8_t *d = reinterpret_cast<uint8_t *>(&hdrs[c]);
    const uint64_t off = blob_off[c];
    for (uint32_t i = 0; i < sizeof(SfqBlobHeader); i++) d[i] = off + i < in_size ? in[off + i] : 0;
}
