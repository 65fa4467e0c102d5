// Two-phase encoder for the gen and qlt streams.
//
// An adaptive range coder is two recurrences glued together: the *model* recurrence (the frequency
// table of a context changes only when that context is visited) and the *coder* recurrence
// (low/range change with every symbol).  An encoder knows every context from the input alone, so
// only the second one is a chain over the whole stream.  The kernels here split them:
//
//   phase 1 (parallel)  for every symbol, the (cumFreq, freq, totFreq) triple its model hands to
//                       RCoder::Encode (coder.hpp:66), in stream order:
//       gen   k_gen_model    one warp per chunk walks the bases 32 at a time; lanes whose contexts
//                            differ update their table slots independently, lanes sharing a context
//                            are replayed in lane order (base2_ranger.hpp:74-84, gens.cpp:138-159)
//       qlt   k_qlt_keys     context of every quality (qlts.hpp:52-74) + per-context histogram
//             k_qlt_scan     exclusive scan -> one segment per visited context
//             k_qlt_scatter  stable counting sort of the positions by context
//             k_qlt_model    one thread per context replays its Log64Ranger (log64_ranger.hpp:69-112)
//                            over the context's positions with the whole model in shared memory
//   phase 2 (chain)     k_rc_encode   one thread per chunk-stream feeds the triples to the coder:
//                       a divide, two multiplies and the renormalisation loop per symbol.
//
// The bytes are those of the single-pass coder (same triples, same order); what changes is that the
// per-symbol table work no longer sits on a 438 k-step dependency chain per chunk.
#pragma once
#include "sfq_streams.cuh"

// ---------------------------------------------------------------------------- packed coding steps
// gen: cum (10 bit) | freq (8 bit) << 10 | tot (10 bit) << 18          (tot <= 4 * 255)
// qlt: cum (22 bit) | freq (17 bit) << 22 | tot (22 bit) << 39 | escape-follows << 61
//      (tot <= 64 * 65472 + 64 < 2^22, freq <= 65473; power_ranger escape steps use the same packing:
//       tot <= 256 * 32736 + 256 < 2^24 does NOT fit 22 bits, so escape steps use SFQ_ESTEP_* below)
SFQ_HD uint32_t sfq_gstep_pack(uint32_t cum, uint32_t f, uint32_t tot) { return cum | (f << 10) | (tot << 18); }
SFQ_HD uint64_t sfq_qstep_pack(uint32_t cum, uint32_t f, uint32_t tot, bool esc) {
    return (uint64_t)cum | ((uint64_t)f << 22) | ((uint64_t)tot << 39) | ((uint64_t)(esc ? 1u : 0u) << 61);
}
// escape (256-symbol model) steps: cum (24) | freq (16) << 24 | tot (24) << 40
SFQ_HD uint64_t sfq_estep_pack(uint32_t cum, uint32_t f, uint32_t tot) {
    return (uint64_t)cum | ((uint64_t)f << 24) | ((uint64_t)tot << 40);
}

// A coder stand-in that records the triples instead of coding them (phase 1 of the escape model).
struct SfqStepSink {
    uint64_t *out;
    SFQ_HD void encode(uint32_t cum, uint32_t freq, uint32_t tot) { *out++ = sfq_estep_pack(cum, freq, tot); }
};

// Where the intermediate arrays of a wave's chunk live.
struct SfqEnc2Chunk {
    uint64_t goff;          // first entry of the chunk in gsteps
    uint64_t qoff;          // first entry in qkey / qb / sorted / qsteps
    uint64_t eoff;          // first entry in esorted / esteps
    uint32_t ecap;          // entries reserved there
    uint32_t pad;
};
#define SFQ_Q_NCTX   65536u
#define SFQ_Q_CNT    65544u           // counters per chunk: 65536 contexts + [65536] = escapes, padded
#define SFQ_SEG_BIG  2048u            // segments at least this long are scheduled first
struct SfqSeg { uint32_t chunk, start, count, flags; };
struct SfqU2 { uint32_t x, y; };                              // (uint2 outside nvcc)      // flags bit 0: escape segment
struct SfqEnc2Ws {
    uint32_t *gsteps;
    uint16_t *qkey;
    uint8_t  *qb;
    uint32_t *sorted;       // pos (24 bit) | symbol << 24, grouped by context, position order inside
    uint64_t *qsteps;
    uint32_t *cnt;          // SFQ_Q_CNT per chunk: histogram, then cursors
    uint32_t *esorted;
    uint64_t *esteps;
    SfqSeg   *segs;         // small segments grow from the front, big ones from the back
    uint64_t  seg_cap;
    uint32_t *ctr;          // [0] small segments, [1] big segments, [2] next segment to hand out, [4] next (chunk, partition) of k_gen_replay
    SfqU2    *gbins;        // gen: (position | symbol << 24, context) grouped by context partition, stream order inside
    uint32_t *gcnt;         // gen: per chunk, gp partition sizes -> end offsets
    uint32_t  gp_bits;      // gen: log2 of the partitions per chunk
};

// ---------------------------------------------------------------------------- Log64Ranger, replay form
// One context's model with O(1) slot lookup and two-level cumulative sums, sized for shared memory
// (log64_ranger.hpp:36-96).  `W` = distance in 32-bit words between consecutive fields of one model
// (models of neighbouring threads are interleaved at an odd stride to spread the banks).
//   words 0..31   freq[64] as u16 pairs
//   words 32..47  syms[64] bytes
//   words 48..63  pos[64] bytes (inverse of syms)
//   words 64..71  gsum[8]  (sum of freq over slots 8g..8g+7)
//   word  72      total;   word 73  count
#define SFQ_L64R_WORDS 75u
struct SfqL64Replay {
    uint32_t *w;
    SFQ_HD uint16_t *freq() const { return reinterpret_cast<uint16_t *>(w); }
    SFQ_HD uint8_t *syms() const { return reinterpret_cast<uint8_t *>(w + 32); }
    SFQ_HD uint8_t *pos() const { return reinterpret_cast<uint8_t *>(w + 48); }
    SFQ_HD uint32_t *gsum() const { return w + 64; }
    SFQ_HD void reset() {
        for (uint32_t k = 0; k < 32; k++) w[k] = 0;
        for (uint32_t k = 0; k < 16; k++) { const uint32_t b = 4 * k; const uint32_t v = b | ((b + 1) << 8) | ((b + 2) << 16) | ((b + 3) << 24); w[32 + k] = v; w[48 + k] = v; }
        for (uint32_t k = 64; k < 74; k++) w[k] = 0;
    }
    // Log64Ranger::put + update_freq for symbol `sym` (< 64); returns the packed coding step.
    SFQ_HD uint64_t step(uint32_t sym, bool esc) {
        uint16_t *fr = freq();
        const uint32_t i = pos()[sym], g = i >> 3, k = i & 7u;
        uint32_t f = fr[i], tot = w[72];
        uint32_t sumf = 0;
        const uint32_t *gs = gsum();
#pragma unroll
        for (uint32_t j = 0; j < 7; j++) sumf += j < g ? gs[j] : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 7; j++) sumf += j < k ? (uint32_t)fr[(i & ~7u) + j] : 0u;
        const uint64_t st = sfq_qstep_pack(sumf + i, f + 1u, tot + 64u, esc);
        // update_freq (log64_ranger.hpp:69-87)
        if (f > 65472u - 6u) {
            if (i == 0 && f + 20u > tot) return st;
            tot = 0;
            for (uint32_t gg = 0; gg < 8; gg++) {
                uint32_t s = 0;
                for (uint32_t j = 0; j < 8; j++) { const uint32_t h = fr[8 * gg + j] >> 1; fr[8 * gg + j] = (uint16_t)h; s += h; }
                gsum()[gg] = s;
                tot += s;
            }
            f = fr[i];
        }
        f += 6u;
        fr[i] = (uint16_t)f;
        gsum()[g] += 6u;
        w[72] = tot + 6u;
        if (i != 0) {
            const uint32_t count = (w[73] + 1u) & 0xffu;
            w[73] = count;
            if ((count & 0xfu) == 0) {
                const uint32_t fp = fr[i - 1];
                if (f > fp) {                               // down_level(): swap slots i and i-1
                    uint8_t *sy = syms(), *ps = pos();
                    const uint8_t sp = sy[i - 1];
                    sy[i - 1] = (uint8_t)sym; sy[i] = sp;
                    ps[sym] = (uint8_t)(i - 1); ps[sp] = (uint8_t)i;
                    fr[i - 1] = (uint16_t)f; fr[i] = (uint16_t)fp;
                    if (k == 0) { gsum()[g - 1] += f - fp; gsum()[g] -= f - fp; }
                }
            }
        }
        return st;
    }
};

// Context used to code position i of a record, in closed form: b1, b2, b3 = the symbols at i-1, i-2, i-3
// (0 before the record starts), delta = 5 + the drops max(0, b[j-1] - b[j]) summed over j <= i-1
// (qlts.hpp:52-74, qlts.cpp:109-134; the reference's q1/q2 ping-pong is just "previous, one before").
SFQ_HD uint32_t sfq_q_ctx_closed(int level, uint32_t i, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t delta) {
    if (level <= 1) return (b1 | (b2 << 6)) & 0xFFFu;
    if (level == 2) return (b1 | (b2 << 6) | (b3 << 12)) & 0xFFFFu;
    if (i == 0) return 0;
    const uint32_t d = delta >> 3;
    return (b1 | ((b2 < b3 ? b3 : b2) << 6) | ((uint32_t)(b2 == b3) << 12) | ((d < 7u ? d : 7u) << 13)) & 0xFFFFu;
}

// ---------------------------------------------------------------------------- phase 2: the coder chain
SFQ_HDN void sfq_rc_gen_chunk(const uint32_t *steps, uint32_t n, uint8_t *out, uint32_t cap, uint32_t *size_out, bool *ovf) {
    SfqEnc rc;
    rc.start(out, cap);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t s = steps[i];
        rc.encode(s & 1023u, (s >> 10) & 255u, s >> 18);
    }
    rc.finish();
    *size_out = rc.out.n;
    *ovf = rc.out.overflow();
}
SFQ_HDN void sfq_rc_qlt_chunk(const uint64_t *steps, const uint64_t *esteps, uint32_t n, uint8_t *out, uint32_t cap,
                              uint32_t *size_out, bool *ovf) {
    SfqEnc rc;
    rc.start(out, cap);
    uint32_t e = 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t s = steps[i];
        rc.encode((uint32_t)s & 0x3fffffu, (uint32_t)(s >> 22) & 0x1ffffu, (uint32_t)(s >> 39) & 0x3fffffu);
        if (s >> 61) {                                                          // qlts.cpp:120-125
            const uint64_t x = esteps[e++];
            rc.encode((uint32_t)x & 0xffffffu, (uint32_t)(x >> 24) & 0xffffu, (uint32_t)(x >> 40));
        }
    }
    rc.finish();
    *size_out = rc.out.n;
    *ovf = rc.out.overflow();
}

#if defined(__CUDACC__)
// ============================================================================ gen, phase 1
// Context of a quality / base position needs the symbols before it in the record: windows of 32
// positions, history carried between windows in `prev`.
__device__ __forceinline__ uint32_t sfq_ldcg32(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ uint64_t sfq_ldcg64(const uint64_t *p) { return __ldcg(reinterpret_cast<const unsigned long long *>(p)); }

// SFQ_GM_BATCH = windows whose text and table slots are requested together; MINB = CTAs per SM the register
// allocation must allow (8 -> 64 registers: every chunk of a 4 741-chunk wave resident at once)
template <int SFQ_GM_BATCH, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_gen_model(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, SfqChunkMeta *metas,
            SfqArena *arenas, uint8_t *arena_buf, SfqWorkspace ws, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c,
            int level, uint32_t nchunks) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (c >= nchunks) return;
    SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    SfqArena *ar = &arenas[c];
    uint32_t *pwpool = ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS;
    SfqXSave xns, xnn;
    xns.init(pwpool, SFQ_X_NS, arena_buf + ar->off[SFQ_S_GEN_NS], ar->cap[SFQ_S_GEN_NS]);
    xnn.init(pwpool, SFQ_X_NN, arena_buf + ar->off[SFQ_S_GEN_NN], ar->cap[SFQ_S_GEN_NN]);
    const bool dense = level <= 1;
    uint64_t *slots = reinterpret_cast<uint64_t *>(ws.gtab + (size_t)c * ws.gtab_stride);
    uint32_t *dtab = reinterpret_cast<uint32_t *>(slots);
    const uint32_t hbits = ws.hbits, hmask = (1u << hbits) - 1u;
    const uint32_t mask = sfq_gen_mask(level);
    const uint32_t solid = meta->solid, nrec = meta->nrec;
    const uint64_t line0 = meta->line0;
    uint32_t *gsteps = e2.gsteps + e2c[c].goff;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;

    uint32_t gbase = 0;                  // bases coded so far (g_genofs_count before this window)
    uint64_t ns_index = 0, nn_index = 0;
    uint32_t n_byte = 0, used = 0;
    uint32_t status = SFQ_OK, status_arg = 0;

    for (uint32_t r = 0; r < nrec && status == SFQ_OK; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, line0, r, solid);
        if (r + 2 < nrec && lane < 4) sfq_prefetch(text + ls[line0 + 4ull * (r + 2)] + 128u * lane);   // the record after next
        uint32_t prev = 0x007616c7u;                                           // gens.cpp:139
        for (uint32_t w0 = 0; w0 < v.llen && status == SFQ_OK; w0 += 32 * SFQ_GM_BATCH) {
            // ---- the batch's text: all loads in flight together
            uint32_t gg[SFQ_GM_BATCH], qq[SFQ_GM_BATCH], ctxs[SFQ_GM_BATCH];
#pragma unroll
            for (int j = 0; j < SFQ_GM_BATCH; j++) {
                const uint32_t i = w0 + 32u * j + lane;
                const bool active = i < v.llen;
                gg[j] = active ? v.seq[i] : (uint32_t)'A';
                qq[j] = active ? (i < v.qlen ? v.qual[i] : 40u) : (uint32_t)'I';            // gens.cpp:153
            }
            // ---- contexts of the batch (2 bits per base, newest lowest) and a prefetch of their home slots
#pragma unroll
            for (int j = 0; j < SFQ_GM_BATCH; j++) {
                if (w0 + 32u * j < v.llen) {
                    uint32_t n = sfq_gencode((uint8_t)gg[j]);
                    if (n > 3u) n = 0;
                    uint32_t h = n, t;
                    t = __shfl_up_sync(FULL, h, 1); if (lane >= 1) h |= t << 2;
                    t = __shfl_up_sync(FULL, h, 2); if (lane >= 2) h |= t << 4;
                    t = __shfl_up_sync(FULL, h, 4); if (lane >= 4) h |= t << 8;
                    t = __shfl_up_sync(FULL, h, 8); if (lane >= 8) h |= t << 16;
                    t = __shfl_up_sync(FULL, h, 1);
                    const uint32_t ctx = ((lane < 16 ? prev << (2 * lane) : 0u) | (lane ? t : 0u)) & mask;
                    prev = __shfl_sync(FULL, h, 31);
                    ctxs[j] = ctx;
                    if (w0 + 32u * j + lane < v.llen) {
                        if (dense) sfq_prefetch(dtab + ctx);
                        else sfq_prefetch(slots + ((ctx * 2654435761u) >> (32 - hbits)));
                    }
                }
            }
            // ---- window by window: exceptions, table slots, replay
#pragma unroll
            for (int j = 0; j < SFQ_GM_BATCH; j++) {
                const uint32_t wj = w0 + 32u * j;
                if (wj < v.llen && status == SFQ_OK) {
                    const bool active = wj + lane < v.llen;
                    const uint32_t g = gg[j], ctx = ctxs[j];
                    uint32_t n = sfq_gencode((uint8_t)g);
                    const bool bad_n = n == 4u;
                    const unsigned errb = __ballot_sync(FULL, n > 4u);
                    if (bad_n) n = 0;
                    const bool bad_q = qq[j] == (uint32_t)'!';
                    unsigned excb = __ballot_sync(FULL, active && (bad_n || bad_q));
                    if (errb) {
                        status = SFQ_E_BASE; status_arg = __shfl_sync(FULL, g, __ffs(errb) - 1);
                    }
                    // exception lists, in base order (gens.cpp:91-114); rare
                    while (excb && status == SFQ_OK) {
                        const int b = __ffs(excb) - 1;
                        excb &= excb - 1;
                        const bool bn = __shfl_sync(FULL, (int)bad_n, b) != 0, bq = __shfl_sync(FULL, (int)bad_q, b) != 0;
                        const uint32_t gb = __shfl_sync(FULL, g, b);
                        const uint64_t genofs = (uint64_t)gbase + (uint32_t)b + 1u;
                        if (!bn) { if (lane == 0) xnn.put(genofs - nn_index); nn_index = genofs; }
                        else {
                            if (!n_byte) n_byte = gb;
                            if (gb != n_byte) { status = SFQ_E_NBYTE; status_arg = gb; break; }
                            if (!bq) { if (lane == 0) xns.put(genofs - ns_index); ns_index = genofs; }
                        }
                    }
                    if (status == SFQ_OK) {
                        // lanes sharing a context form a group; its first lane finds (or claims) the slot
                        const unsigned peers = __match_any_sync(FULL, active ? ctx : 0xffffffffu);
                        const uint32_t rank = __popc(peers & lt), gsize = __popc(peers);
                        uint32_t slot = 0, fv = 0x03030303u;
                        bool need = active && rank == 0, full = false;
                        if (dense) { if (need) { slot = ctx; fv = sfq_ldcg32(dtab + ctx) ^ 0x03030303u; } }
                        else {
                            const uint32_t key = ctx + 1u;
                            uint32_t hh = (ctx * 2654435761u) >> (32 - hbits), probes = 0;
                            while (__any_sync(FULL, need)) {
                                bool claim = false;
                                if (need) {
                                    const uint64_t k = sfq_ldcg64(slots + hh);
                                    if ((uint32_t)(k >> 32) == key) { slot = hh; fv = (uint32_t)k; need = false; }
                                    else if (k == 0) claim = true;
                                    else { hh = (hh + 1u) & hmask; if (++probes > hmask) { full = true; need = false; } }
                                }
                                // two contexts of the window may want the same empty slot: the lower lane takes it
                                const unsigned cl = __match_any_sync(FULL, claim ? hh : (0x80000000u | lane));
                                if (claim) {
                                    if ((cl & lt) == 0) { slots[hh] = ((uint64_t)key << 32) | 0x03030303ull; slot = hh; used++; need = false; }
                                    else { hh = (hh + 1u) & hmask; probes++; }
                                }
                                __syncwarp();
                            }
                        }
                        if (__any_sync(FULL, full)) status = SFQ_E_TABLE;
                        const int leader = __ffs(peers) - 1;
                        slot = __shfl_sync(FULL, slot, leader);
                        fv = __shfl_sync(FULL, fv, leader);
                        // replay the group in lane order
                        const uint32_t maxrank = __reduce_max_sync(FULL, active ? rank : 0u);
                        uint32_t step = 0;
                        for (uint32_t rr = 0; rr <= maxrank; rr++) {
                            if (rank == rr) {
                                const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
                                const uint32_t cum = (n > 0 ? f0 : 0) + (n > 1 ? f1 : 0) + (n > 2 ? f2 : 0);
                                step = sfq_gstep_pack(cum, (fv >> (8 * n)) & 0xffu, f0 + f1 + f2 + f3);
                                fv = sfq_b2_update(fv, n);
                            }
                            if (rr < maxrank) {
                                const int src = __fns(peers, 0, rr + 1);       // lane holding rank rr of my group
                                const uint32_t nv = __shfl_sync(FULL, fv, src < 0 ? 0 : src);
                                if (rank > rr) fv = nv;
                            }
                        }
                        if (active && status == SFQ_OK) {
                            gsteps[gbase + lane] = step;
                            if (rank + 1 == gsize) {
                                if (dense) dtab[slot] = fv ^ 0x03030303u;
                                else slots[slot] = ((uint64_t)(ctx + 1u) << 32) | fv;
                            }
                        }
                        gbase += min(32u, v.llen - wj);
                        __syncwarp();
                    }
                }
            }
        }
    }
    used = __reduce_add_sync(FULL, used);
    if (lane == 0) {
        bool ovf = false;
        ar->size[SFQ_S_GEN_NS] = xns.close(ovf);
        ar->size[SFQ_S_GEN_NN] = xnn.close(ovf);
        if (status == SFQ_OK && ovf) status = SFQ_E_CAP;
        meta->n_byte = (n_byte && n_byte != 'N') ? (uint8_t)n_byte : 0;        // gens.cpp:103-104
        meta->g_used = used;
        if (status != SFQ_OK && meta->status == SFQ_OK) { meta->status = status; meta->status_arg = status_arg; }
    }
}

// ============================================================================ gen, phase 1, partitioned form
// k_gen_model above keeps a chunk's 4-symbol models in a hash table in global memory: one random 8-byte slot
// per base, i.e. two DRAM sectors per base and 8 MiB of table per 1 MiB chunk, walked by ONE warp per chunk.
// The kernels below bring the models into shared memory instead:
//   k_gen_count    how many bases of the chunk fall into each of GP partitions of the context space
//                  (partition = top bits of a multiplicative hash of the context); order-free, all records in parallel
//   k_gen_scatter  one warp per chunk walks the bases in stream order (exception lists gen.Ns / gen.Nn are
//                  written on the way, gens.cpp:91-114) and appends (position, symbol, context) to the list of
//                  the context's partition - a stable partition, 8 bytes per base.  Lists start on 32-byte
//                  boundaries and are staged four entries at a time in shared memory, so what goes to DRAM
//                  are whole sectors
//   k_gen_replay   one warp per (chunk, partition): the partition's few hundred contexts live in a 2048-slot hash
//                  table in the warp's shared memory; the warp replays the list 32 entries at a time - every lane
//                  finds or claims its context's slot by itself (compare-and-swap on the key), windows in which two
//                  lanes meet in one slot take the ordered path (entries of one context in list order,
//                  base2_ranger.hpp:74-84) - and writes each base's coding step to its stream position.
// A context belongs to exactly one partition and a partition's list keeps stream order, so every model sees its
// symbols in the reference's order: the steps, hence the bytes, are those of the single-pass coder.
#define SFQ_GP_TARGET   1024u          // most bases per partition, on average, the partition count aims for (sfq_gen_gp_bits)
#define SFQ_GP_SLOTS    2048u          // hash slots per replay warp (8 bytes each) + one owner byte per slot
#define SFQ_GP_SMEM_MAX 4096u          // the scatter warp keeps cursors and staging rows in shared memory up to this many partitions
#define SFQ_GR_WARPS    12             // replay warps per CTA (12 x 18 KB)
#define SFQ_GR_BATCH    8u             // (chunk, partition) items a replay warp takes per trip to the work counter
#define SFQ_GR_SMEM     (SFQ_GR_WARPS * (SFQ_GP_SLOTS * 8u + SFQ_GP_SLOTS))
#define SFQ_GC_RECS     128u           // records per CTA of k_gen_count
__device__ __forceinline__ uint32_t sfq_gp_of(uint32_t ctx, uint32_t gp_bits) { return gp_bits ? (ctx * 2654435761u) >> (32u - gp_bits) : 0u; }
__device__ __forceinline__ uint32_t sfq_gp_slot(uint32_t ctx, uint32_t gp_bits) { return ((ctx * 2654435761u) >> (21u - gp_bits)) & (SFQ_GP_SLOTS - 1u); }   // the 11 bits below the partition's (gp_bits <= 14)
__host__ __device__ __forceinline__ uint64_t sfq_gbins_off(uint64_t goff, uint32_t c_in_wave, uint32_t gp_bits) { return goff + ((uint64_t)c_in_wave << (gp_bits + 2u)); }   // lists are padded to 4 entries

// Contexts of a window of 32 bases (2 bits per base, newest lowest; gens.cpp:139-147): n = this lane's code,
// prev = context carried in from the bases before the window (updated for the next window).
__device__ __forceinline__ uint32_t sfq_gen_window_ctx(uint32_t n, uint32_t lane, uint32_t &prev, uint32_t mask) {
    const unsigned FULL = 0xffffffffu;
    uint32_t h = n, t;
    t = __shfl_up_sync(FULL, h, 1); if (lane >= 1) h |= t << 2;
    t = __shfl_up_sync(FULL, h, 2); if (lane >= 2) h |= t << 4;
    t = __shfl_up_sync(FULL, h, 4); if (lane >= 4) h |= t << 8;
    t = __shfl_up_sync(FULL, h, 8); if (lane >= 8) h |= t << 16;
    t = __shfl_up_sync(FULL, h, 1);
    const uint32_t ctx = ((lane < 16 ? prev << (2 * lane) : 0u) | (lane ? t : 0u)) & mask;
    prev = __shfl_sync(FULL, h, 31);
    return ctx;
}
// Lanes holding the same `bits`-bit value as this lane (one ballot per bit: the cost does not depend on how many
// different values the warp holds, unlike __match_any_sync, whose hardware loop takes one turn per distinct value).
__device__ __forceinline__ unsigned sfq_match_bits(uint32_t v, uint32_t bits, unsigned among) {
    unsigned peers = among;
    for (uint32_t b = 0; b < bits; b++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (v >> b) & 1u);
        peers &= ((v >> b) & 1u) ? bal : ~bal;
    }
    return peers;
}

// Key of a base: context (26 bits) | symbol << 26 | "is an N" << 28 | "quality is '!'" << 29 | which N byte << 30 (N, n, .).
// gcnt[c * gp + p] += bases of chunk c whose context lies in partition p.  Order-free: a CTA takes SFQ_GC_RECS records
// of one chunk (blockIdx.y), a warp one record at a time.  The keys lie in the chunk's step array, which k_gen_replay
// only fills after k_gen_part has read them.
#define SFQ_GK_N     (1u << 28)
#define SFQ_GK_BADQ  (1u << 29)
__global__ void __launch_bounds__(128)
k_gen_keys(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, SfqChunkMeta *metas, const uint32_t *__restrict__ rec_boff,
           SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, int level, uint32_t nchunks) {
    const uint32_t c = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (c >= nchunks) return;
    SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    const uint32_t nrec = meta->nrec;
    if (blockIdx.x * SFQ_GC_RECS >= nrec) return;
    const unsigned FULL = 0xffffffffu;
    const uint32_t gp_bits = e2.gp_bits;
    uint32_t *cnt = e2.gcnt + ((size_t)c << gp_bits);
    uint32_t *gkey = e2.gsteps + e2c[c].goff;
    const uint32_t solid = meta->solid, mask = sfq_gen_mask(level);
    const uint64_t line0 = meta->line0;
    const uint32_t *boffs = rec_boff + line0 / 4;
    const uint32_t r_end = min(nrec, (blockIdx.x + 1u) * SFQ_GC_RECS);
    for (uint32_t r = blockIdx.x * SFQ_GC_RECS + warp; r < r_end; r += 4) {
        const SfqRecView v = sfq_rec_view(text, ls, line0, r, solid);
        const uint32_t bbase = boffs[r];
        uint32_t prev = 0x007616c7u;                                           // gens.cpp:139
        for (uint32_t w0 = 0; w0 < v.llen; w0 += 32) {
            const uint32_t i = w0 + lane;
            const bool active = i < v.llen;
            const uint32_t g = active ? v.seq[i] : (uint32_t)'A';
            const uint32_t q = active ? (i < v.qlen ? v.qual[i] : 40u) : (uint32_t)'I';        // gens.cpp:153
            uint32_t n = sfq_gencode((uint8_t)g);
            const unsigned errb = __ballot_sync(FULL, active && n > 4u);
            if (errb) {                                                        // unexpected genome char (gens.cpp:125-126): the earliest one is reported
                if (active && n > 4u) atomicMax(&meta->status_arg, ((0xffffffu - (bbase + i)) << 8) | g);    // (largest = earliest position)
                if (lane == 0) atomicCAS(&meta->status, (uint32_t)SFQ_OK, (uint32_t)SFQ_E_BASE);
            }
            const bool bad_n = n == 4u;
            if (n > 3u) n = 0;
            const uint32_t ctx = sfq_gen_window_ctx(n, lane, prev, mask);
            if (active) {
                gkey[bbase + i] = ctx | (n << 26) | (bad_n ? SFQ_GK_N : 0u) | (q == (uint32_t)'!' ? SFQ_GK_BADQ : 0u) |
                                  ((g == (uint32_t)'n' ? 1u : g == (uint32_t)'.' ? 2u : 0u) << 30);
                atomicAdd(cnt + sfq_gp_of(ctx, gp_bits), 1u);
            }
        }
    }
}

// One warp per chunk walks the keys in stream order: exception lists (gens.cpp:91-114), then (position | symbol << 24,
// context) appended to the list of the context's partition.  gcnt holds the partition sizes on entry and each
// partition's END offset on exit (list p starts at the end of list p-1 rounded up to 4 entries).
__global__ void __launch_bounds__(32)
k_gen_part(SfqChunkMeta *metas, SfqArena *arenas, uint8_t *arena_buf, SfqWorkspace ws, SfqEnc2Ws e2,
           const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    extern __shared__ uint4 sc_smem[];                     // gp staging rows of 4 entries (32 bytes), then gp cursors
    const uint32_t c = blockIdx.x;
    const uint32_t lane = threadIdx.x;
    if (c >= nchunks) return;
    SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    SfqArena *ar = &arenas[c];
    uint32_t *pwpool = ws.pw + (size_t)c * SFQ_PW_PER_CHUNK * SFQ_PW_WORDS;
    SfqXSave xns, xnn;
    xns.init(pwpool, SFQ_X_NS, arena_buf + ar->off[SFQ_S_GEN_NS], ar->cap[SFQ_S_GEN_NS]);
    xnn.init(pwpool, SFQ_X_NN, arena_buf + ar->off[SFQ_S_GEN_NN], ar->cap[SFQ_S_GEN_NN]);
    const uint32_t gp_bits = e2.gp_bits, gp = 1u << gp_bits;
    uint32_t *gc = e2.gcnt + (size_t)c * gp;
    const bool staged = gp <= SFQ_GP_SMEM_MAX;             // beyond that (chunks of > 4 M bases): cursors in global memory, entries written one by one
    uint2 *stage = reinterpret_cast<uint2 *>(sc_smem);
    uint32_t *cur = staged ? reinterpret_cast<uint32_t *>(sc_smem + 2 * (size_t)gp) : gc;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    {   // exclusive prefix of the padded partition sizes -> cursors (lane k owns a block of consecutive partitions)
        const uint32_t per = (gp + 31u) / 32u, lo = min(gp, lane * per), hi = min(gp, lo + per);
        uint32_t s = 0;
        for (uint32_t k = lo; k < hi; k++) s += (gc[k] + 3u) & ~3u;
        uint32_t inc = s, t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { t = __shfl_up_sync(FULL, inc, d); if (lane >= (uint32_t)d) inc += t; }
        uint32_t run = inc - s;
        for (uint32_t k = lo; k < hi; k++) { const uint32_t v = (gc[k] + 3u) & ~3u; cur[k] = run; run += v; }
        __syncwarp();
    }
    const uint32_t *gkey = e2.gsteps + e2c[c].goff;
    uint2 *bins = reinterpret_cast<uint2 *>(e2.gbins) + sfq_gbins_off(e2c[c].goff, c, gp_bits);
    const uint32_t nb = meta->nbases;
    uint64_t ns_index = 0, nn_index = 0;
    uint32_t n_byte = 0;
    uint32_t status = SFQ_OK, status_arg = 0;
    constexpr int B = 8;
    for (uint32_t p0 = 0; p0 < nb && status == SFQ_OK; p0 += 32 * B) {
        uint32_t keys[B];
#pragma unroll
        for (int j = 0; j < B; j++) { const uint32_t pos = p0 + 32u * j + lane; keys[j] = pos < nb ? gkey[pos] : 0u; }
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint32_t w0 = p0 + 32u * j, pos = w0 + lane;
            if (w0 < nb && status == SFQ_OK) {
                const bool active = pos < nb;
                const uint32_t key = keys[j], ctx = key & 0x3ffffffu, n = (key >> 26) & 3u;
                unsigned excb = __ballot_sync(FULL, active && (key & (SFQ_GK_N | SFQ_GK_BADQ)));
                while (excb && status == SFQ_OK) {                             // exception lists, in base order (gens.cpp:91-114); rare
                    const int b = __ffs(excb) - 1;
                    excb &= excb - 1;
                    const uint32_t kb = __shfl_sync(FULL, key, b);
                    const bool bn = (kb & SFQ_GK_N) != 0, bq = (kb & SFQ_GK_BADQ) != 0;
                    const uint32_t gb = (kb >> 30) == 1u ? (uint32_t)'n' : (kb >> 30) == 2u ? (uint32_t)'.' : (uint32_t)'N';
                    const uint64_t genofs = (uint64_t)w0 + (uint32_t)b + 1u;
                    if (!bn) { if (lane == 0) xnn.put(genofs - nn_index); nn_index = genofs; }
                    else {
                        if (!n_byte) n_byte = gb;
                        if (gb != n_byte) { status = SFQ_E_NBYTE; status_arg = gb; break; }
                        if (!bq) { if (lane == 0) xns.put(genofs - ns_index); ns_index = genofs; }
                    }
                }
                if (status == SFQ_OK) {
                    const unsigned am = __ballot_sync(FULL, active);
                    const uint32_t p = sfq_gp_of(ctx, gp_bits);
                    unsigned peers = sfq_match_bits(p, gp_bits, am);
                    if (!active) peers = 1u << lane;
                    const uint32_t rank = __popc(peers & lt), k = __popc(peers);
                    uint32_t at0 = 0;
                    if (active && rank == 0) { at0 = cur[p]; cur[p] = at0 + k; }
                    at0 = __shfl_sync(FULL, at0, __ffs(peers) - 1);
                    const uint32_t at = at0 + rank;
                    const uint2 e = make_uint2(pos | (n << 24), ctx);
                    if (!staged) { if (active) bins[at] = e; }
                    else {
                        // the rows (4 entries = one sector) this window's group of partition p touches: entries of its first
                        // row join what earlier windows staged; rows it fills alone go out directly (four lanes, one sector);
                        // an unfinished last row is staged once the first one has left
                        const uint32_t row = at >> 2, first = at0 >> 2, last = (at0 + k - 1u) >> 2;
                        const bool last_done = ((at0 + k - 1u) & 3u) == 3u;
                        const bool in_first = active && row == first;
                        const bool direct = active && row != first && (row != last || last_done);
                        const bool in_last = active && row != first && !direct;
                        if (in_first) stage[4u * p + (at & 3u)] = e;
                        if (direct) bins[at] = e;
                        __syncwarp();
                        if (in_first && (at & 3u) == 3u) {
                            const uint4 *rowp = reinterpret_cast<const uint4 *>(stage + 4u * p);
                            uint4 *dst = reinterpret_cast<uint4 *>(bins + (at - 3u));
                            dst[0] = rowp[0]; dst[1] = rowp[1];
                        }
                        __syncwarp();
                        if (in_last) stage[4u * p + (at & 3u)] = e;
                    }
                    __syncwarp();
                }
            }
        }
    }
    if (staged) {
        for (uint32_t k = lane; k < gp; k += 32) {                             // unfinished rows, and the end offsets for k_gen_replay
            const uint32_t endk = cur[k];
            for (uint32_t q = endk & ~3u; q < endk; q++) bins[q] = stage[4u * k + (q & 3u)];
            gc[k] = endk;
        }
    }
    if (lane == 0) {
        bool ovf = false;
        ar->size[SFQ_S_GEN_NS] = xns.close(ovf);
        ar->size[SFQ_S_GEN_NN] = xnn.close(ovf);
        if (status == SFQ_OK && ovf) status = SFQ_E_CAP;
        meta->n_byte = (n_byte && n_byte != 'N') ? (uint8_t)n_byte : 0;        // gens.cpp:103-104
        if (status != SFQ_OK && meta->status == SFQ_OK) { meta->status = status; meta->status_arg = status_arg; }
    }
}

// One warp per (chunk, partition), handed out through e2.ctr[4] in chunk-major order, SFQ_GR_BATCH at a time (the
// partitions of a chunk run close together in time, so the chunk's step array fills in L2 before it goes out to DRAM).
__global__ void __launch_bounds__(32 * SFQ_GR_WARPS)
k_gen_replay(SfqChunkMeta *metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    extern __shared__ uint4 gr_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint2 *tab = reinterpret_cast<uint2 *>(gr_smem) + (size_t)warp * SFQ_GP_SLOTS;                       // (key, freq) slots
    uint8_t *owner = reinterpret_cast<uint8_t *>(reinterpret_cast<uint2 *>(gr_smem) + (size_t)SFQ_GR_WARPS * SFQ_GP_SLOTS) + (size_t)warp * SFQ_GP_SLOTS;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t gp_bits = e2.gp_bits, gp = 1u << gp_bits;
    const uint64_t nitems = (uint64_t)nchunks << gp_bits;
    for (;;) {
        uint32_t item0 = 0;
        if (lane == 0) item0 = atomicAdd(e2.ctr + 4, SFQ_GR_BATCH);
        item0 = __shfl_sync(FULL, item0, 0);
        if (item0 >= nitems) break;
        for (uint32_t item = item0; item < item0 + SFQ_GR_BATCH && item < nitems; item++) {
            const uint32_t c = item >> gp_bits, p = item & (gp - 1u);
            if (metas[c].status != SFQ_OK) continue;
            const uint32_t *ends = e2.gcnt + (size_t)c * gp;
            const uint32_t start = p ? (ends[p - 1] + 3u) & ~3u : 0u, end = ends[p];
            if (end <= start) continue;
            {   // empty table
                uint4 *t4 = reinterpret_cast<uint4 *>(tab);
#pragma unroll 4
                for (uint32_t k = lane; k < SFQ_GP_SLOTS / 2; k += 32) t4[k] = make_uint4(0, 0, 0, 0);
                __syncwarp();
            }
            const uint2 *bins = reinterpret_cast<const uint2 *>(e2.gbins) + sfq_gbins_off(e2c[c].goff, c, gp_bits);
            uint32_t *gsteps = e2.gsteps + e2c[c].goff;
            uint32_t used = 0;
            bool full = false;
            uint2 nxt = start + lane < end ? bins[start + lane] : make_uint2(0, 0);
            for (uint32_t w0 = start; w0 < end; w0 += 32) {
                const bool active = w0 + lane < end;
                const uint2 e = nxt;
                if (w0 + 32 + lane < end) nxt = bins[w0 + 32 + lane];             // the next window's entries travel during this one
                const uint32_t ctx = e.y, n = (e.x >> 24) & 3u, pos = e.x & 0xffffffu;
                const uint32_t key = ctx + 1u;
                // every lane finds or claims the slot of its context by itself
                uint32_t slot = sfq_gp_slot(ctx, gp_bits);
                bool fresh = false;
                if (active) {
                    for (uint32_t probes = 0;; probes++) {
                        uint32_t k = reinterpret_cast<volatile uint2 *>(tab)[slot].x;
                        if (k == 0u) { k = atomicCAS(&tab[slot].x, 0u, key); if (k == 0u) { fresh = true; break; } }
                        if (k == key) break;
                        if (probes >= SFQ_GP_SLOTS) { full = true; break; }
                        slot = (slot + 1u) & (SFQ_GP_SLOTS - 1u);
                    }
                    owner[slot] = (uint8_t)lane;
                }
                used += (uint32_t)__popc(__ballot_sync(FULL, fresh));             // (warp-uniform count of occupied slots)
                if (used > SFQ_GP_SLOTS - 64u) full = true;                         // nearly full: probes get long; rerun with more partitions
                __syncwarp();
                const bool meet = active && owner[slot] != (uint8_t)lane;         // somebody else of this window is in my slot: same context
                uint32_t step;
                if (!__any_sync(FULL, meet)) {
                    uint32_t fv = fresh ? 0x03030303u : tab[slot].y;
                    const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
                    const uint32_t cum = (n > 0 ? f0 : 0) + (n > 1 ? f1 : 0) + (n > 2 ? f2 : 0);
                    step = sfq_gstep_pack(cum, (fv >> (8 * n)) & 0xffu, f0 + f1 + f2 + f3);
                    if (active) tab[slot].y = sfq_b2_update(fv, n);
                } else {
                    // ordered path: lanes sharing a slot (= a context) form a group, replayed in lane order
                    const unsigned am = __ballot_sync(FULL, active);
                    const unsigned peers = active ? __match_any_sync(am, slot) : 1u << lane;
                    const uint32_t rank = __popc(peers & lt), gsize = __popc(peers);
                    // the group's start state: the table's, or the initial one if the slot was claimed in this window
                    const unsigned freshb = __ballot_sync(FULL, fresh);
                    uint32_t fv = (freshb & peers) ? 0x03030303u : (active ? tab[slot].y : 0u);
                    const uint32_t maxrank = __reduce_max_sync(FULL, active ? rank : 0u);
                    step = 0;
                    for (uint32_t rr = 0; rr <= maxrank; rr++) {
                        if (rank == rr) {
                            const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
                            const uint32_t cum = (n > 0 ? f0 : 0) + (n > 1 ? f1 : 0) + (n > 2 ? f2 : 0);
                            step = sfq_gstep_pack(cum, (fv >> (8 * n)) & 0xffu, f0 + f1 + f2 + f3);
                            fv = sfq_b2_update(fv, n);
                        }
                        if (rr < maxrank) {
                            const int src = __fns(peers, 0, rr + 1);               // lane holding rank rr of my group
                            const uint32_t nv = __shfl_sync(FULL, fv, src < 0 ? 0 : src);
                            if (rank > rr) fv = nv;
                        }
                    }
                    if (active && rank + 1 == gsize) tab[slot].y = fv;
                }
                if (active) gsteps[pos] = step;
                __syncwarp();
            }
            if (__any_sync(FULL, full)) { if (lane == 0) atomicCAS(&metas[c].status, (uint32_t)SFQ_OK, (uint32_t)SFQ_E_TABLE); }
            else if (lane == 0) atomicAdd(&metas[c].g_used, used);
        }
    }
}

// ============================================================================ qlt, phase 1
// Context of position i from b[i-1], b[i-2], b[i-3] and the running sum of drops (qlts.hpp:52-74,
// qlts.cpp:109-134).  Lanes = 32 consecutive positions of a record.
struct SfqQWin {
    uint32_t p1, p2, p3, dsum;           // carries from the previous window of the record
    __device__ __forceinline__ void reset() { p1 = p2 = p3 = 0; dsum = 0; }
    // b = this lane's symbol (0 for inactive lanes; they are the tail of the record)
    __device__ __forceinline__ uint32_t ctx(uint32_t b, uint32_t lane, uint32_t i, int level) {
        const unsigned FULL = 0xffffffffu;
        uint32_t b1 = __shfl_up_sync(FULL, b, 1), b2 = __shfl_up_sync(FULL, b, 2), b3 = __shfl_up_sync(FULL, b, 3);
        if (lane < 1) b1 = p1;
        if (lane < 2) b2 = lane == 0 ? p2 : p1;
        if (lane < 3) b3 = lane == 0 ? p3 : lane == 1 ? p2 : p1;
        uint32_t e = b2 > b1 ? b2 - b1 : 0u, t;          // drop paid when position i-1 was coded
        t = __shfl_up_sync(FULL, e, 1);  if (lane >= 1)  e += t;
        t = __shfl_up_sync(FULL, e, 2);  if (lane >= 2)  e += t;
        t = __shfl_up_sync(FULL, e, 4);  if (lane >= 4)  e += t;
        t = __shfl_up_sync(FULL, e, 8);  if (lane >= 8)  e += t;
        t = __shfl_up_sync(FULL, e, 16); if (lane >= 16) e += t;
        const uint32_t delta = 5u + dsum + e;
        // carries for the next window: positions 31, 30, 29 of this one (lane 0 of the next window adds
        // the drop of position 31 itself)
        const uint32_t n1 = __shfl_sync(FULL, b, 31), n2 = __shfl_sync(FULL, b, 30), n3 = __shfl_sync(FULL, b, 29);
        const uint32_t e31 = __shfl_sync(FULL, e, 31);
        dsum += e31;
        p1 = n1; p2 = n2; p3 = n3;
        return sfq_q_ctx_closed(level, i, b1, b2, b3, delta);
    }
};

// Contexts and the per-context histogram do not depend on any order: a CTA takes SFQ_QK_RECS records of
// one chunk (blockIdx.y), a warp one record at a time; counts go out as fire-and-forget atomics.
#define SFQ_QK_RECS 64
__global__ void __launch_bounds__(128)
k_qlt_keys(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, SfqChunkMeta *metas,
           const uint32_t *__restrict__ rec_qoff, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, int level, uint32_t nchunks) {
    const uint32_t c = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (c >= nchunks) return;
    const SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    const unsigned FULL = 0xffffffffu;
    const uint32_t solid = meta->solid, nrec = meta->nrec;
    const uint64_t line0 = meta->line0;
    const uint32_t *qoffs = rec_qoff + line0 / 4;
    uint16_t *qkey = e2.qkey + e2c[c].qoff;
    uint8_t *qb = e2.qb + e2c[c].qoff;
    uint32_t *cnt = e2.cnt + (size_t)c * SFQ_Q_CNT;
    const uint32_t r_end = min(nrec, (blockIdx.x + 1u) * SFQ_QK_RECS);
    uint32_t nesc = 0;
    for (uint32_t r = blockIdx.x * SFQ_QK_RECS + warp; r < r_end; r += 4) {
        const SfqRecView v = sfq_rec_view(text, ls, line0, r, solid);
        const uint32_t qbase = qoffs[r];
        SfqQWin win; win.reset();
        for (uint32_t w0 = 0; w0 < v.qlen; w0 += 32) {
            const uint32_t i = w0 + lane;
            const bool active = i < v.qlen;
            const uint32_t b = active ? (uint32_t)(uint8_t)(v.qual[i] - '!') : 0u;
            const uint32_t ctx = win.ctx(b, lane, i, level);
            nesc += __popc(__ballot_sync(FULL, active && b >= 63u));
            const unsigned peers = __match_any_sync(FULL, active ? ctx : 0xffffffffu);
            if (active) {
                qkey[qbase + i] = (uint16_t)ctx;
                qb[qbase + i] = (uint8_t)b;
                if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(cnt + ctx, (uint32_t)__popc(peers));
            }
        }
    }
    if (lane == 0 && nesc) atomicAdd(cnt + SFQ_Q_NCTX, nesc);
}

// One CTA per chunk: histogram -> cursors (in place) + the segment list.
__global__ void __launch_bounds__(256)
k_qlt_scan(SfqChunkMeta *metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    const uint32_t c = blockIdx.x;
    SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    uint32_t *cnt = e2.cnt + (size_t)c * SFQ_Q_CNT;
    // thread t owns contexts t, t + 256, ...: coalesced, and the order of the segments does not matter
    constexpr uint32_t PER = SFQ_Q_NCTX / 256;
    uint32_t s = 0, ks = 0, kb = 0;
    for (uint32_t k = 0; k < PER; k++) {
        const uint32_t v = cnt[k * 256 + threadIdx.x];
        s += v;
        if (v >= SFQ_SEG_BIG) kb++; else if (v) ks++;
    }
    __shared__ uint32_t sh[3][256];
    __shared__ uint32_t base[2];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = ks; sh[2][threadIdx.x] = kb;
    __syncthreads();
    if (threadIdx.x < 3) {               // three tiny serial scans, one thread each
        uint32_t run = 0;
        uint32_t *a = sh[threadIdx.x];
        for (int k = 0; k < 256; k++) { const uint32_t v = a[k]; a[k] = run; run += v; }
        const uint32_t nesc = cnt[SFQ_Q_NCTX];
        if (threadIdx.x == 1) base[0] = atomicAdd(e2.ctr + 0, run);
        if (threadIdx.x == 2) base[1] = atomicAdd(e2.ctr + 1, run + (nesc ? 1u : 0u)) + (nesc ? 1u : 0u);
        if (threadIdx.x == 1) meta->q_used = run;           // + big ones, added below
        if (threadIdx.x == 0) {
            meta->extra_hi = nesc;
            if (nesc > e2c[c].ecap && meta->status == SFQ_OK) meta->status = SFQ_E_CAP;
        }
    }
    __syncthreads();
    if (threadIdx.x == 255) atomicAdd(&meta->q_used, sh[2][255] + kb);
    if (threadIdx.x == 0) {
        const uint32_t nesc = cnt[SFQ_Q_NCTX];
        if (nesc) {                      // the escape model's segment goes first among the big ones
            SfqSeg sg; sg.chunk = c; sg.start = 0; sg.count = nesc; sg.flags = 1;
            e2.segs[e2.seg_cap - 1 - (base[1] - 1u)] = sg;
        }
        cnt[SFQ_Q_NCTX] = 0;             // becomes the escape cursor
    }
    uint32_t run = sh[0][threadIdx.x], is = base[0] + sh[1][threadIdx.x], ib = base[1] + sh[2][threadIdx.x];
    for (uint32_t k = 0; k < PER; k++) {
        const uint32_t v = cnt[k * 256 + threadIdx.x];
        cnt[k * 256 + threadIdx.x] = run;
        if (v) {
            SfqSeg sg; sg.chunk = c; sg.start = run; sg.count = v; sg.flags = 0;
            if (v >= SFQ_SEG_BIG) e2.segs[e2.seg_cap - 1 - ib++] = sg; else e2.segs[is++] = sg;
        }
        run += v;
    }
}

// Stable counting sort: positions must reach their segment in order, so one warp walks the chunk;
// keys and cursors of eight windows are requested together to keep the walk off the HBM latency.
__global__ void __launch_bounds__(128)
k_qlt_scatter(const SfqChunkMeta *__restrict__ metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (c >= nchunks) return;
    const SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint16_t *qkey = e2.qkey + e2c[c].qoff;
    const uint8_t *qb = e2.qb + e2c[c].qoff;
    uint32_t *sorted = e2.sorted + e2c[c].qoff;
    uint32_t *esorted = e2.esorted + e2c[c].eoff;
    uint32_t *cur = e2.cnt + (size_t)c * SFQ_Q_CNT;
    const uint32_t n = meta->nquals;
    uint32_t ecur = 0;
    constexpr int B = 8;
    for (uint32_t p0 = 0; p0 < n; p0 += 32 * B) {
        uint32_t keys[B], bs[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint32_t p = p0 + 32u * j + lane;
            keys[j] = p < n ? (uint32_t)qkey[p] : 0xffffffffu;
            bs[j] = p < n ? (uint32_t)qb[p] : 0u;
        }
#pragma unroll
        for (int j = 0; j < B; j++) if (keys[j] != 0xffffffffu) sfq_prefetch(cur + keys[j]);
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint32_t p = p0 + 32u * j + lane;
            if (p0 + 32u * j < n) {
                const bool active = p < n;
                const uint32_t ctx = keys[j], b = bs[j];
                const unsigned peers = __match_any_sync(FULL, ctx);
                const uint32_t rank = __popc(peers & lt);
                uint32_t at = 0;
                if (active && rank == 0) { at = sfq_ldcg32(cur + ctx); cur[ctx] = at + __popc(peers); }
                at = __shfl_sync(FULL, at, __ffs(peers) - 1) + rank;
                if (active) sorted[at] = p | ((b < 63u ? b : 63u) << 24);
                const unsigned eb = __ballot_sync(FULL, active && b >= 63u);
                if (active && b >= 63u) esorted[ecur + __popc(eb & lt)] = p | (b << 24);
                ecur += __popc(eb);
                __syncwarp();
            }
        }
    }
}

// The same stable counting sort in two levels, each of which writes few, short-lived streams:
//   k_qlt_part1   one warp per chunk walks the positions in order and appends (position | symbol << 24, context >> 8)
//                 to the list of the context's LOW byte - 256 lists whose starts k_qlt_scan has already laid out
//                 (thread t of the scan owns contexts t, t + 256, ...: their segments are contiguous).  Entries are
//                 staged four at a time in shared memory, so DRAM sees whole 32-byte rows.  The lists live in the
//                 memory of the coding steps (8 bytes per quality), which are only written after level 2 has read them
//   k_qlt_part2   one warp per (chunk, low byte): 256 cursors (the segments of the contexts with this low byte) in
//                 shared memory; the list's entries reach their final places inside the list's own region of `sorted`
// Against k_qlt_scatter's one pass over 65 536 cursors in global memory (a DRAM round trip and two partial sectors
// per quality) this moves 23 bytes per quality in streams.
#define SFQ_QP_BATCH 8u
__global__ void __launch_bounds__(32)
k_qlt_part1(const SfqChunkMeta *__restrict__ metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    __shared__ uint2 stage[256 * 4];
    __shared__ uint32_t cur[256], lo[256];
    const uint32_t c = blockIdx.x, lane = threadIdx.x;
    if (c >= nchunks) return;
    const SfqChunkMeta *meta = &metas[c];
    if (meta->status != SFQ_OK) return;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint16_t *qkey = e2.qkey + e2c[c].qoff;
    const uint8_t *qb = e2.qb + e2c[c].qoff;
    uint2 *tmp = reinterpret_cast<uint2 *>(e2.qsteps + e2c[c].qoff);
    uint32_t *esorted = e2.esorted + e2c[c].eoff;
    const uint32_t *cnt = e2.cnt + (size_t)c * SFQ_Q_CNT;
    for (uint32_t t = lane; t < 256; t += 32) { const uint32_t v = cnt[t]; cur[t] = v; lo[t] = v; }     // start of context t = start of list t
    __syncwarp();
    const uint32_t n = meta->nquals;
    uint32_t ecur = 0;
    constexpr int B = 8;
    for (uint32_t p0 = 0; p0 < n; p0 += 32 * B) {
        uint32_t keys[B], bs[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint32_t p = p0 + 32u * j + lane;
            keys[j] = p < n ? (uint32_t)qkey[p] : 0u;
            bs[j] = p < n ? (uint32_t)qb[p] : 0u;
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
            const uint32_t p = p0 + 32u * j + lane;
            if (p0 + 32u * j < n) {
                const bool active = p < n;
                const uint32_t ctx = keys[j], b = bs[j], t = ctx & 255u;
                const unsigned am = __ballot_sync(FULL, active);
                unsigned peers = sfq_match_bits(t, 8, am);
                if (!active) peers = 1u << lane;
                const uint32_t rank = __popc(peers & lt), k = __popc(peers);
                uint32_t at0 = 0;
                if (active && rank == 0) { at0 = cur[t]; cur[t] = at0 + k; }
                at0 = __shfl_sync(FULL, at0, __ffs(peers) - 1);
                const uint32_t at = at0 + rank;
                const uint2 e = make_uint2(p | ((b < 63u ? b : 63u) << 24), ctx >> 8);
                // rows of four entries (one sector), as in k_gen_scatter; a list may start or end inside a row
                const uint32_t row = at >> 2, first = at0 >> 2, last = (at0 + k - 1u) >> 2;
                const bool last_done = ((at0 + k - 1u) & 3u) == 3u;
                const bool in_first = active && row == first;
                const bool direct = active && row != first && (row != last || last_done);
                const bool in_last = active && row != first && !direct;
                if (in_first) stage[4u * t + (at & 3u)] = e;
                if (direct) tmp[at] = e;
                __syncwarp();
                if (in_first && (at & 3u) == 3u) {
                    const uint32_t from = max(at - 3u, lo[t]);
                    if (from == at - 3u) {
                        const uint4 *rowp = reinterpret_cast<const uint4 *>(stage + 4u * t);
                        uint4 *dst = reinterpret_cast<uint4 *>(tmp + from);
                        dst[0] = rowp[0]; dst[1] = rowp[1];
                    } else for (uint32_t q = from; q <= at; q++) tmp[q] = stage[4u * t + (q & 3u)];
                }
                __syncwarp();
                if (in_last) stage[4u * t + (at & 3u)] = e;
                const unsigned eb = __ballot_sync(FULL, active && b >= 63u);          // escapes, in stream order (qlts.cpp:120-125)
                if (active && b >= 63u) esorted[ecur + __popc(eb & lt)] = p | (b << 24);
                ecur += __popc(eb);
                __syncwarp();
            }
        }
    }
    for (uint32_t t = lane; t < 256; t += 32) {                                       // unfinished rows
        const uint32_t endt = cur[t];
        for (uint32_t q = max(endt & ~3u, lo[t]); q < endt; q++) tmp[q] = stage[4u * t + (q & 3u)];
    }
}

__global__ void __launch_bounds__(256)
k_qlt_part2(const SfqChunkMeta *__restrict__ metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    __shared__ uint32_t curs[8][256];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *cur = curs[warp];
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t nitems = (uint64_t)nchunks << 8;
    for (;;) {
        uint32_t item0 = 0;
        if (lane == 0) item0 = atomicAdd(e2.ctr + 5, SFQ_QP_BATCH);
        item0 = __shfl_sync(FULL, item0, 0);
        if (item0 >= nitems) break;
        for (uint32_t item = item0; item < item0 + SFQ_QP_BATCH && item < nitems; item++) {
            const uint32_t c = item >> 8, t = item & 255u;
            if (metas[c].status != SFQ_OK) continue;
            const uint32_t *cnt = e2.cnt + (size_t)c * SFQ_Q_CNT;
            // list t = the segments of contexts t, t + 256, ...: it starts where context t starts and ends where list t + 1
            // starts (the last one at the chunk's quality count)
            const uint32_t start = cnt[t], end = t == 255u ? metas[c].nquals : cnt[t + 1u];
            if (end <= start) continue;
            __syncwarp();
            for (uint32_t k = lane; k < 256; k += 32) cur[k] = cnt[(k << 8) + t];
            __syncwarp();
            const uint2 *tmp = reinterpret_cast<const uint2 *>(e2.qsteps + e2c[c].qoff);
            uint32_t *sorted = e2.sorted + e2c[c].qoff;
            uint2 nxt = start + lane < end ? tmp[start + lane] : make_uint2(0, 0);
            for (uint32_t w0 = start; w0 < end; w0 += 32) {
                const bool active = w0 + lane < end;
                const uint2 e = nxt;
                if (w0 + 32 + lane < end) nxt = tmp[w0 + 32 + lane];
                const uint32_t k = e.y & 255u;
                const unsigned am = __ballot_sync(FULL, active);
                unsigned peers = sfq_match_bits(k, 8, am);
                if (!active) peers = 1u << lane;
                const uint32_t rank = __popc(peers & lt);
                uint32_t at = 0;
                if (active && rank == 0) { at = cur[k]; cur[k] = at + (uint32_t)__popc(peers); }
                at = __shfl_sync(FULL, at, __ffs(peers) - 1) + rank;
                if (active) sorted[at] = e.x;
                __syncwarp();
            }
        }
    }
}

// One thread per context segment (handed out dynamically, long ones first): the whole model lives in
// shared memory for the lifetime of the segment.
#define SFQ_QM_THREADS 128
__global__ void __launch_bounds__(SFQ_QM_THREADS)
k_qlt_model(const SfqChunkMeta *__restrict__ metas, SfqWorkspace ws, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t chunk0) {
    __shared__ uint32_t models[SFQ_QM_THREADS * SFQ_L64R_WORDS];
    SfqL64Replay m;
    m.w = models + threadIdx.x * SFQ_L64R_WORDS;
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nsmall = e2.ctr[0], nbig = e2.ctr[1], total = nsmall + nbig;
    bool have = false, done = false;
    const uint32_t *src = nullptr;
    uint64_t *dst = nullptr;
    uint32_t p = 0, end = 0;
    (void)chunk0;
    for (;;) {
        const unsigned need = __ballot_sync(FULL, !have && !done);
        if (need) {
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(e2.ctr + 2, (uint32_t)__popc(need));
            first = __shfl_sync(FULL, first, 0);
            if (!have && !done) {
                const uint32_t j = first + __popc(need & ((1u << lane) - 1u));
                if (j >= total) done = true;
                else {
                    const SfqSeg sg = j < nbig ? e2.segs[e2.seg_cap - 1 - j] : e2.segs[j - nbig];
                    const SfqEnc2Chunk ec = e2c[sg.chunk];
                    if (sg.flags & 1u) {
                        // escape symbols of the chunk through its 256-symbol model (qlts.cpp:124, power_ranger.hpp:93-106)
                        if (metas[sg.chunk].status == SFQ_OK) {
                            SfqPower ex; ex.m = ws.pw + ((size_t)sg.chunk * SFQ_PW_PER_CHUNK + SFQ_PW_QEX) * SFQ_PW_WORDS;
                            SfqStepSink sink; sink.out = e2.esteps + ec.eoff;
                            const uint32_t *es = e2.esorted + ec.eoff;
                            for (uint32_t k = 0; k < sg.count; k++) ex.put(sink, es[k] >> 24);
                        }
                    } else if (metas[sg.chunk].status == SFQ_OK) {     // (a chunk that overflowed its escape list was not scattered)
                        m.reset();
                        src = e2.sorted + ec.qoff; dst = e2.qsteps + ec.qoff;
                        p = sg.start; end = sg.start + sg.count;
                        have = true;
                    }
                }
            }
        }
        if (__all_sync(FULL, done)) break;
        if (have) {
            const uint32_t n = min(end - p, 8u);
            for (uint32_t k = 0; k < n; k++) {
                const uint32_t e = src[p + k];
                const uint32_t b = e >> 24;
                dst[e & 0xffffffu] = m.step(b, false);
            }
            p += n;
            if (p == end) have = false;
        }
    }
}

// The escape flag of a position is known from the symbol alone; folding it in here keeps k_qlt_model's
// step() free of it.
__global__ void __launch_bounds__(256)
k_qlt_mark_escapes(const SfqChunkMeta *__restrict__ metas, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    const uint32_t c = blockIdx.x;
    if (c >= nchunks || metas[c].status != SFQ_OK || metas[c].extra_hi == 0) return;
    const uint32_t *es = e2.esorted + e2c[c].eoff;
    uint64_t *qs = e2.qsteps + e2c[c].qoff;
    for (uint32_t k = threadIdx.x; k < metas[c].extra_hi; k += blockDim.x) qs[es[k] & 0xffffffu] |= 1ull << 61;
}

// ============================================================================ phase 2
// One thread per chunk-stream; kind 0 = gen, 1 = qlt.  The steps of a stream are read 16 bytes at a
// time through a ring of four registers, four loads ahead of the coder (the stores of the coder's
// output may alias, so the compiler will not hoist loads by itself), and range / totFreq is a
// multiply by a reciprocal: a 1021-entry table in shared memory for the 4-symbol model
// (totFreq <= 1020), computed off the dependency chain for the 64-symbol model.
__device__ __forceinline__ uint32_t sfq_div_by(uint32_t n, uint32_t d, uint32_t inv) {   // inv = floor(2^32 / d), d >= 2
    const uint32_t q = __umulhi(n, inv);
    return q + ((n - q * d) >= d ? 1u : 0u);
}
__device__ __forceinline__ uint32_t sfq_recip32(uint32_t d) {                            // floor(2^32 / d), d >= 2
    const uint32_t q = 0xffffffffu / d;
    return q + ((0xffffffffu - q * d) == d - 1u ? 1u : 0u);
}
// The quality chain with four lanes per chunk-stream (eight streams per warp).  What is NOT part of the chain - fetching a
// step, taking it apart, floor(2^32 / totFreq) (a real division: two thirds of the instructions of the one-lane form) - is done
// for four consecutive steps at once, one per lane; the chain itself then takes each step's fields out of its lane with four
// shuffles.  All four lanes run the coder (same state, same bytes: no divergence inside a group), only lane 0 stores.
__global__ void __launch_bounds__(32)
k_rc_encode_q4(SfqChunkMeta *metas, SfqArena *arenas, uint8_t *arena_buf, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c, uint32_t nchunks) {
    const unsigned FULL = 0xffffffffu;
    const uint32_t g = threadIdx.x >> 2, sub = threadIdx.x & 3u;
    const uint32_t c = blockIdx.x * 8u + g;
    const bool on = c < nchunks && metas[c].status == SFQ_OK;
    SfqEnc rc;
    rc.reset();
    uint32_t n = 0;
    const uint64_t *steps = e2.qsteps, *es = e2.esteps;
    if (on) {
        rc.start(arena_buf + arenas[c].off[SFQ_S_QLT], arenas[c].cap[SFQ_S_QLT]);
        n = metas[c].nquals;
        steps = e2.qsteps + e2c[c].qoff;
        es = e2.esteps + e2c[c].eoff;
    }
    rc.out.mute = !on || sub != 0u;
    uint32_t nmax = n;
    nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 4)); nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 8)); nmax = max(nmax, __shfl_xor_sync(FULL, nmax, 16));
    uint64_t cur = steps[sub], nxt = steps[4u + sub];            // (the arrays are padded: reading up to 8 steps past a stream's end is safe)
    for (uint32_t i = 0; i < nmax; i += 4u) {
        const uint64_t s = cur;
        cur = nxt;
        nxt = i < n ? steps[i + 8u + sub] : 0ull;                  // (a short stream beside long ones must not read on behind its own padding)
        uint32_t tot = (uint32_t)(s >> 39) & 0x3fffffu;
        if (tot < 64u) tot = 64u;                                 // tot >= 64 in every real step; the clamp only guards the read-ahead padding
        const uint32_t inv = sfq_recip32(tot);
        const uint32_t cum = (uint32_t)s & 0x3fffffu;
        const uint32_t fe = ((uint32_t)(s >> 22) & 0x1ffffu) | ((uint32_t)(s >> 61) << 31);      // freq | escape-follows << 31
#pragma unroll
        for (uint32_t k = 0; k < 4u; k++) {
            const uint32_t ck = __shfl_sync(FULL, cum, k, 4), fk = __shfl_sync(FULL, fe, k, 4);
            const uint32_t tk = __shfl_sync(FULL, tot, k, 4), ik = __shfl_sync(FULL, inv, k, 4);
            if (i + k < n) {
                rc.encode_scaled(ck, fk & 0x1ffffu, sfq_div_by(rc.range, tk, ik));
                if (fk >> 31) {                                   // qlts.cpp:120-125
                    const uint64_t x = *es++;
                    rc.encode((uint32_t)x & 0xffffffu, (uint32_t)(x >> 24) & 0xffffu, (uint32_t)(x >> 40));
                }
            }
        }
    }
    if (on && sub == 0u) {
        rc.finish();
        arenas[c].size[SFQ_S_QLT] = rc.out.n;
        if (rc.out.overflow()) atomicCAS(&metas[c].status, (uint32_t)SFQ_OK, (uint32_t)SFQ_E_CAP);
    }
}

// (Up to SFQ_RC_MAXW warps per CTA may share the base chain's 4 KB reciprocal table - SFQ_RC_WARPS; measured: no effect on the
// kernels running beside it, so the launch keeps one-warp CTAs.)
#define SFQ_RC_MAXW 4
template <int KIND>
__global__ void __launch_bounds__(32 * SFQ_RC_MAXW)
k_rc_encode(SfqChunkMeta *metas, SfqArena *arenas, uint8_t *arena_buf, SfqEnc2Ws e2, const SfqEnc2Chunk *__restrict__ e2c,
            uint32_t nchunks, uint32_t lanes) {
    __shared__ uint32_t lut[KIND == 0 ? 1024 : 1];
    if (KIND == 0) {
        for (uint32_t t = threadIdx.x; t < 1024; t += blockDim.x) lut[t] = t < 2 ? 0xffffffffu : sfq_recip32(t);
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * lanes + lane;
    if (lane >= lanes || c >= nchunks) return;
    SfqChunkMeta *m = &metas[c];
    if (m->status != SFQ_OK) return;
    SfqArena *ar = &arenas[c];
    const int sid = KIND == 0 ? SFQ_S_GEN : SFQ_S_QLT;
    SfqEnc rc;
    rc.start(arena_buf + ar->off[sid], ar->cap[sid]);
    if (KIND == 0) {
        const uint32_t n = m->nbases;
        const uint4 *p = reinterpret_cast<const uint4 *>(e2.gsteps + e2c[c].goff);
        uint4 b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3];
        auto block = [&](const uint4 &b, uint32_t first) {
            const uint32_t s[4] = {b.x, b.y, b.z, b.w};
            uint32_t inv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) inv[j] = lut[(s[j] >> 18) & 1023u];     // (the read-ahead past the stream's end sees arbitrary words)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (first + j < n) {
                    const uint32_t tot = (s[j] >> 18) & 1023u;
                    rc.encode_scaled(s[j] & 1023u, (s[j] >> 10) & 255u, sfq_div_by(rc.range, tot, inv[j]));
                }
        };
        for (uint32_t i = 0; i < n; i += 16) {
            const uint32_t k = i >> 2;
            block(b0, i);      b0 = p[k + 4];
            block(b1, i + 4);  b1 = p[k + 5];
            block(b2, i + 8);  b2 = p[k + 6];
            block(b3, i + 12); b3 = p[k + 7];
        }
    } else {
        const uint32_t n = m->nquals;
        const uint4 *p = reinterpret_cast<const uint4 *>(e2.qsteps + e2c[c].qoff);
        const uint64_t *es = e2.esteps + e2c[c].eoff;
        uint4 b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3];
        auto block = [&](const uint4 &b, uint32_t first) {
            const uint64_t s[2] = {(uint64_t)b.x | ((uint64_t)b.y << 32), (uint64_t)b.z | ((uint64_t)b.w << 32)};
            uint32_t tot[2], inv[2];
#pragma unroll
            for (int j = 0; j < 2; j++) { tot[j] = (uint32_t)(s[j] >> 39) & 0x3fffffu; if (tot[j] < 64u) tot[j] = 64u; inv[j] = sfq_recip32(tot[j]); }   // tot >= 64 in every real step; the clamp only guards the read-ahead padding
#pragma unroll
            for (int j = 0; j < 2; j++)
                if (first + j < n) {
                    rc.encode_scaled((uint32_t)s[j] & 0x3fffffu, (uint32_t)(s[j] >> 22) & 0x1ffffu, sfq_div_by(rc.range, tot[j], inv[j]));
                    if (s[j] >> 61) {                                           // qlts.cpp:120-125
                        const uint64_t x = *es++;
                        rc.encode((uint32_t)x & 0xffffffu, (uint32_t)(x >> 24) & 0xffffu, (uint32_t)(x >> 40));
                    }
                }
        };
        for (uint32_t i = 0; i < n; i += 8) {
            const uint32_t k = i >> 1;
            block(b0, i);     b0 = p[k + 4];
            block(b1, i + 2); b1 = p[k + 5];
            block(b2, i + 4); b2 = p[k + 6];
            block(b3, i + 6); b3 = p[k + 7];
        }
    }
    rc.finish();
    ar->size[sid] = rc.out.n;
    if (rc.out.overflow()) atomicCAS(&m->status, (uint32_t)SFQ_OK, (uint32_t)SFQ_E_CAP);
}
#endif  // __CUDACC__
