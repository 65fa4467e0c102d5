// Base stream decoder, warp-converged form: 32 chunks per warp, one lane per chunk, no divergent inner loops.
//
// Same arithmetic and the same table as sfq_gen_decode_loop (sfq_streams.cuh: buckets of four slots, line chosen
// by the grandparent context, bucket copied to shared memory one base ahead, line prefetched two ahead,
// range/tot by reciprocal table, symbol search by compare) - what changes is the control flow: the
// thread-per-chunk kernel lets every lane run its own record loop and its own renormalisation loop, so a
// warp's lanes drift apart and most issue slots carry 1-8 active lanes.  Here one loop drives all 32 lanes;
// the renormalisation shifts its whole bytes in at once (as in sfq_qlt_dec.cuh), record changes and the rare
// cases (bucket overflow, carry guard) are predicated under a warp vote.  A wave of 9 481 chunks is 297
// warps, each instruction carrying 32 chunks.
#pragma once
#include "sfq_qlt_dec.cuh"

#if defined(__CUDACC__)

__global__ void __launch_bounds__(32)
k_gen_decode32(const uint8_t *__restrict__ in, const SfqDecChunk *__restrict__ dc, SfqChunkMeta *metas, SfqWorkspace ws,
               SfqRecTables t, uint8_t *bases, uint32_t nchunks) {
    __shared__ uint32_t lut[SFQ_B2_LUT];
    __shared__ uint4 cells[64];
    sfq_b2_lut_fill(lut, threadIdx.x, 32);
    __syncwarp();
    const unsigned FULL = 0xffffffffu;
    const uint32_t c = blockIdx.x * 32u + threadIdx.x;
    bool live = c < nchunks;
    if (live) live = metas[c].status == SFQ_OK;
    SfqStage stage;
    stage.cell0 = &cells[threadIdx.x]; stage.cell1 = &cells[32 + threadIdx.x];

    SfqByteSrc src;
    src.p = nullptr; src.end = nullptr; src.word = 0; src.ahead = 0; src.left = 8;
    SfqGenBuckets tab;
    tab.init(ws.gtab, 4, true);
    uint32_t mask = 0, alpha = 0x54474341u, nrec = 0;
    const uint32_t *llen_tab = t.llen;
    const uint64_t *boff_tab = t.boff;
    bool dense = true;
    uint64_t low = 0, code = 0;
    uint32_t range = 0xFFFFFFFFu;
    if (live) {
        const SfqDecChunk &d = dc[c];
        const int level = d.level;
        dense = level <= 1;
        tab.init(ws.gtab + (size_t)c * ws.gtab_stride, ws.hbits, dense);
        tab.ahead2 = ws.gen_ahead2;
        mask = sfq_gen_mask(level);
        alpha = metas[c].solid ? 0x33323130u : 0x54474341u;                         // "0123" / "ACGT", gens.cpp:173-178
        nrec = metas[c].nrec;
        llen_tab = t.llen + d.rec_base; boff_tab = t.boff + d.rec_base;
        src.start(in + d.soff[SFQ_S_GEN], d.ssize[SFQ_S_GEN]);
        for (int k = 0; k < 8; k++) code = (code << 8) | src.next();               // coder.hpp:44-48
    }
    uint32_t *dtab = reinterpret_cast<uint32_t *>(tab.slots);

    uint32_t r = 0, i = 0, llen = 0, last = 0x007616c7u, bk = 0, bkn = 0;
    uint8_t *g = bases;
    uint32_t pj = 4, pkey = 0, pfv = 0;      // slot of the current bucket written after it was requested (4 = none)
    bool tabfull = false;
    // next non-empty record: its first bucket is requested, the line of its second base prefetched
#define SFQ_GD_OPEN(want)                                                                                    \
    {                                                                                                        \
        bool w_ = (want);                                                                                    \
        while (__any_sync(FULL, w_)) {                                                                       \
            if (w_) {                                                                                        \
                if (r >= nrec) { live = false; w_ = false; }                                                 \
                else {                                                                                       \
                    llen = sfq_coded_len(llen_tab[r]);                                                       \
                    if (llen) {                                                                              \
                        i = 0; g = bases + boff_tab[r]; last = 0x007616c7u; pj = 4; w_ = false;               \
                        if (!dense) {                                                                        \
                            bk = tab.home(last & mask);                                                      \
                            stage.request(tab.slots + 4ull * bk);                                            \
                            if (llen > 1) tab.prefetch_line(tab.next_home(last & mask, mask) >> 2);           \
                        }                                                                                    \
                    } else r++;                                                                              \
                }                                                                                            \
            }                                                                                                \
        }                                                                                                    \
    }
    SFQ_GD_OPEN(live)

    while (__any_sync(FULL, live)) {
        const bool act = live;
        const uint32_t ctx = last & mask;
        uint32_t fv = 0x03030303u;
        uint64_t *slot = nullptr;
        bool ovf = false;
        if (act) {
            if (dense) fv = dtab[ctx] ^ 0x03030303u;
            else {
                uint32_t k[4], v[4];
                stage.collect(tab.slots + 4ull * bk, k, v);
                // the bucket was requested before the previous base stored its slot: if that store went into
                // this very bucket, the copy is one update behind
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) if (q == pj) { k[q] = pkey; v[q] = pfv; }
                uint32_t j; bool hit;
                if (sfq_bucket_pick(k, v, ctx + 1u, j, fv, hit)) { slot = tab.slots + 4ull * bk + j; tab.used += hit ? 0u : 1u; }
                else ovf = true;
            }
        }
        if (__any_sync(FULL, ovf)) {                  // home bucket full of other contexts: walk on (rare)
            if (ovf) {
                uint32_t b2 = bk, j; bool hit;
                for (uint32_t probes = 0; !slot && probes < tab.nb; probes++) {
                    b2 = b2 + 1u == tab.nb ? 0u : b2 + 1u;
                    uint32_t k2[4], v2[4];
                    sfq_ld_bucket(tab.slots + 4ull * b2, k2, v2);
                    if (sfq_bucket_pick(k2, v2, ctx + 1u, j, fv, hit)) { slot = tab.slots + 4ull * b2 + j; tab.used += hit ? 0u : 1u; }
                }
                if (!slot) { tabfull = true; live = false; }
            }
        }
        const bool go = act && live;
        if (go && !dense) {
            // the bucket of the next base does not depend on what this base decodes to: request it now
            bkn = tab.next_home(ctx, mask);
            if (i + 1 < llen) stage.request(tab.slots + 4ull * bkn);
            if (i + 2 < llen) tab.prefetch_line(tab.line_after2(ctx, mask));
        }
        // ---- Base2Ranger::get + RCoder::Decode (base2_ranger.hpp:86-104, coder.hpp:83-102)
        const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
        const uint32_t tot = f0 + f1 + f2 + f3;
        const uint32_t rr = sfq_div_recip(range, tot, lut[tot & (SFQ_B2_LUT - 1u)]);    // GetFreq: range /= tot
        // (a corrupt stream can leave code >= 2^32: every comparison is then true, as with the reference's 64-bit quotient)
        const uint32_t c32 = (uint32_t)(code >> 32) ? 0xffffffffu : (uint32_t)code;
        const uint32_t c1 = f0 * rr, c2 = c1 + f1 * rr, c3 = c2 + f2 * rr;
        const bool g1 = c32 >= c1, g2 = c32 >= c2, g3 = c32 >= c3;
        const uint32_t b = (uint32_t)g1 + (uint32_t)g2 + (uint32_t)g3;
        const uint32_t cr = g3 ? c3 : g2 ? c2 : g1 ? c1 : 0u;
        const uint32_t f = g3 ? f3 : g2 ? f2 : g1 ? f1 : f0;
        if (go) {
            low += cr;
            code -= cr;
            range = rr * f;
            // renormalise: n whole bytes at once; the carry guard (coder.hpp:95-96) can only fire when bits
            // 32..39 of low are all ones - then take the reference's loop
            const uint32_t n = (uint32_t)__clz((int)range) >> 3;
            if (n) {
                if (((uint32_t)(low >> 32) & 0xffu) == 0xffu) {
                    while (range < SFQ_RC_TOP) {
                        if ((low ^ (low + range)) & (0xffULL << 56)) range = (((uint32_t)low) | (SFQ_RC_TOP - 1)) - (uint32_t)low;
                        code = (code << 8) | src.next();
                        range <<= 8;
                        low <<= 8;
                    }
                } else {
                    const uint32_t v = sfq_src_take(src, n);
                    code = (code << (8u * n)) | v;
                    range <<= 8u * n;
                    low <<= 8u * n;
                }
            }
            fv = sfq_b2_update(fv, b);
            if (dense) dtab[ctx] = fv ^ 0x03030303u;
            else {
                *slot = ((uint64_t)(ctx + 1u) << 32) | fv;
                const uint32_t sidx = (uint32_t)(slot - tab.slots);
                pj = (sidx >> 2) == bkn ? (sidx & 3u) : 4u; pkey = ctx + 1u; pfv = fv;
                bk = bkn;
            }
            last = (last << 2) + b;
            g[i] = (uint8_t)(alpha >> (8 * b));       // exceptions: k_gen_exceptions, afterwards
            i++;
        }
        const bool endrec = go && i == llen;
        if (endrec) r++;
        SFQ_GD_OPEN(endrec)
    }
#undef SFQ_GD_OPEN
    if (c < nchunks && tabfull && metas[c].status == SFQ_OK) metas[c].status = SFQ_E_TABLE;
}

#endif  // __CUDACC__
