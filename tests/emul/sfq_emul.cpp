// TEST TOOLING ONLY: compiles the per-chunk coder routines of slimfastq_b200/csrc (the bodies of the
// CUDA kernels) with g++ and runs them one chunk-stream at a time on the CPU, so their logic can be
// checked against the oracle in the GPU-less dev container.  Never linked into the product library.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../slimfastq_b200/csrc/sfq_streams.cuh"
#include "../../slimfastq_b200/csrc/sfq_encode2.cuh"
#include <map>
#include "../../slimfastq_b200/csrc/sfq_plan.cuh"
#include "../../slimfastq_b200/csrc/sfq_container.h"
#include "../../slimfastq_b200/csrc/sfq_layout.h"

static void put_bytes(std::vector<uint8_t> &o, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p; o.insert(o.end(), b, b + n);
}

// the device buffers are 16-byte aligned and padded; give the CPU run the same guarantees
static std::vector<uint8_t> padded(const uint8_t *p, size_t n) {
    std::vector<uint8_t> v(n + 64, 0);
    if (n) memcpy(v.data(), p, n);
    return v;
}

// The two-phase encoder's algorithm, run sequentially: coding steps of every symbol first (gen: table walk;
// qlt: positions grouped by context, one replay model per context, escapes through the 256-symbol model),
// then the coder chain over the steps.  Shares sfq_q_ctx_closed, SfqL64Replay, the step packing and
// sfq_rc_*_chunk with the kernels; the warp choreography itself is checked on the GPU.
static int g_two_phase = 0;
extern "C" void sfq_emul_set_two_phase(int on) { g_two_phase = on; }

static void emul_gen_two_phase(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *m, int level, void *gt, uint32_t hbits,
                               uint8_t *arena, SfqArena *ar) {
    SfqGenTable tab; tab.init(gt, hbits, level <= 1);
    const uint32_t mask = sfq_gen_mask(level);
    std::vector<uint32_t> steps;
    for (uint32_t r = 0; r < m->nrec; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, m->line0, r, m->solid);
        uint32_t last = 0x007616c7u;
        for (uint32_t i = 0; i < v.llen; i++) {
            uint32_t n = sfq_gencode(v.seq[i]);
            if (n > 3) n = 0;
            last &= mask;
            uint32_t fv; const uint32_t slot = tab.find(last, fv);
            const uint32_t f0 = fv & 0xff, f1 = (fv >> 8) & 0xff, f2 = (fv >> 16) & 0xff, f3 = fv >> 24;
            steps.push_back(sfq_gstep_pack((n > 0 ? f0 : 0) + (n > 1 ? f1 : 0) + (n > 2 ? f2 : 0), (fv >> (8 * n)) & 0xff, f0 + f1 + f2 + f3));
            tab.store(slot, last, sfq_b2_update(fv, n));
            last = (last << 2) | n;
        }
    }
    bool ovf = false;
    sfq_rc_gen_chunk(steps.data(), (uint32_t)steps.size(), arena + ar->off[SFQ_S_GEN], ar->cap[SFQ_S_GEN], &ar->size[SFQ_S_GEN], &ovf);
    if (ovf && m->status == SFQ_OK) m->status = SFQ_E_CAP;
}

static void emul_qlt_two_phase(const uint8_t *text, const uint64_t *ls, SfqChunkMeta *m, int level, uint32_t *pw,
                               uint8_t *arena, SfqArena *ar) {
    std::vector<uint32_t> key, sym;
    for (uint32_t r = 0; r < m->nrec; r++) {
        const SfqRecView v = sfq_rec_view(text, ls, m->line0, r, m->solid);
        uint32_t delta = 5;
        for (uint32_t i = 0; i < v.qlen; i++) {
            auto at = [&](int64_t k) -> uint32_t { return k < 0 ? 0u : (uint32_t)(uint8_t)(v.qual[k] - '!'); };
            const uint32_t b1 = at((int64_t)i - 1), b2 = at((int64_t)i - 2), b3 = at((int64_t)i - 3);
            if (i >= 1 && b2 > b1) delta += b2 - b1;
            key.push_back(sfq_q_ctx_closed(level, i, b1, b2, b3, delta));
            sym.push_back(at(i));
        }
    }
    std::map<uint32_t, std::vector<uint32_t>> groups;
    for (uint32_t p = 0; p < key.size(); p++) groups[key[p]].push_back(p);
    std::vector<uint64_t> steps(key.size()), esteps;
    for (auto &g : groups) {
        uint32_t words[SFQ_L64R_WORDS];
        SfqL64Replay mod; mod.w = words; mod.reset();
        for (uint32_t p : g.second) steps[p] = mod.step(sym[p] < 63 ? sym[p] : 63, false);
    }
    uint32_t extra = 0;
    for (uint32_t p = 0; p < key.size(); p++) if (sym[p] >= 63) extra++;
    esteps.resize(extra + 1);
    SfqPower ex; ex.m = pw + (size_t)SFQ_PW_QEX * SFQ_PW_WORDS;
    SfqStepSink sink; sink.out = esteps.data();
    for (uint32_t p = 0; p < key.size(); p++) if (sym[p] >= 63) { ex.put(sink, sym[p]); steps[p] |= 1ull << 61; }
    bool ovf = false;
    sfq_rc_qlt_chunk(steps.data(), esteps.data(), (uint32_t)steps.size(), arena + ar->off[SFQ_S_QLT], ar->cap[SFQ_S_QLT], &ar->size[SFQ_S_QLT], &ovf);
    m->extra_hi = extra;
    if (ovf && m->status == SFQ_OK) m->status = SFQ_E_CAP;
}

static uint64_t g_phase = 0;
extern "C" void sfq_emul_set_chunk_phase(uint64_t phase) { g_phase = phase; }

extern "C" int sfq_emul_compress(const uint8_t *text_in, size_t n, int level, uint64_t chunk_bytes,
                                 uint8_t **out, size_t *out_n, uint32_t *status_out) {
    const uint64_t phase = g_phase;
    g_phase = 0;
    const std::vector<uint8_t> text_copy = padded(text_in, n);
    const uint8_t *text = text_copy.data();
    level = level > 4 ? 4 : level < 1 ? 1 : level;
    *status_out = 0;
    std::vector<uint64_t> ls;
    ls.push_back(0);
    for (size_t i = 0; i < n; i++) if (text[i] == '\n') ls.push_back(i + 1);
    if (n == 0 || text[n - 1] != '\n' || (ls.size() - 1) % 4) { *status_out = SFQ_E_TRUNC; return 1; }
    const uint64_t nrec = (ls.size() - 1) / 4;
    const uint64_t nslots = sfq_slot_count(n, chunk_bytes, phase);
    std::vector<uint8_t> file(sizeof(SfqFileHeader));
    std::vector<uint64_t> index;
    uint64_t out_total = 0;
    for (uint64_t c = 0; c < nslots; c++) {
        uint64_t r0 = sfq_first_record_at(ls.data(), nrec, sfq_slot_target(c, chunk_bytes, phase));
        uint64_t r1 = c + 1 == nslots ? nrec : sfq_first_record_at(ls.data(), nrec, sfq_slot_target(c + 1, chunk_bytes, phase));
        if (r0 == r1) continue;
        SfqChunkMeta m;
        sfq_plan_chunk(text, ls.data(), r0, r1, &m);
        if (m.status) { *status_out = m.status; return 1; }
        for (uint32_t grow = 0;; grow++) {
            SfqArena ar; uint64_t end;
            sfq_arena_layout(&m, grow, 0, &ar, &end);
            std::vector<uint8_t> arena(end);
            uint32_t hbits = sfq_gen_hbits(level, m.nbases, grow);
            void *gt = calloc(1, sfq_gtable_bytes(level, hbits));
            uint32_t *qt = (uint32_t *)calloc(1, sfq_qtable_bytes(level));
            uint32_t *pw = (uint32_t *)calloc(1, sfq_pwpool_bytes());
            m.status = 0;
            sfq_gen_encode_chunk(text, ls.data(), &m, level, gt, hbits, pw, arena.data(), &ar);
            if (g_two_phase) {      // gen stream and qlt stream again, the two-phase way, over fresh tables
                void *gt2 = calloc(1, sfq_gtable_bytes(level, hbits));
                emul_gen_two_phase(text, ls.data(), &m, level, gt2, hbits, arena.data(), &ar);
                free(gt2);
                emul_qlt_two_phase(text, ls.data(), &m, level, pw, arena.data(), &ar);
            } else
            sfq_qlt_encode_chunk(text, ls.data(), &m, level, qt, pw, arena.data(), &ar);
            { static SfqRecScratch scr; sfq_rec_encode_chunk(text, ls.data(), &m, pw, arena.data(), &ar, &scr); }
            free(gt); free(qt); free(pw);
            if ((m.status == SFQ_E_CAP || m.status == SFQ_E_TABLE) && grow < 6) continue;
            if (m.status) { *status_out = m.status; return 1; }
            SfqBlobHeader b; memset(&b, 0, sizeof b);
            b.magic = SFQ_BLOB_MAGIC; b.level = level; b.text_len = m.text_len; b.out_len = m.out_len; b.nrec = m.nrec;
            b.nbases = m.nbases; b.nquals = m.nquals; b.hdr_bytes = m.hdr_bytes; b.llen = m.llen;
            b.solid = m.solid; b.two_id = m.two_id; b.n_byte = m.n_byte; b.extra_hi = m.extra_hi;
            const uint64_t fl = m.line0 + 4ull * m.first_coded;
            b.rec_first_len = m.first_coded < m.nrec ? (uint32_t)(ls[fl + 1] - ls[fl] - 2) : 0; b.g_used = m.g_used;
            b.nbig = m.nbig; b.big_bases = m.big_bases; b.big_quals = m.big_quals; b.big_hdr = m.big_hdr;
            for (int k = 0; k < SFQ_NSTREAMS; k++) b.ssize[k] = ar.size[k];
            index.push_back(file.size());
            out_total += m.out_len;
            put_bytes(file, &b, sizeof b);
            put_bytes(file, text + ls[fl] + 1, b.rec_first_len);
            for (int k = 0; k < SFQ_NSTREAMS; k++) put_bytes(file, arena.data() + ar.off[k], ar.size[k]);
            break;
        }
    }
    SfqFileHeader h;
    sfq_file_header_init(&h, level, n, index.size(), chunk_bytes, file.size(), out_total);
    memcpy(file.data(), &h, sizeof h);
    put_bytes(file, index.data(), index.size() * 8);
    *out = (uint8_t *)malloc(file.size());
    memcpy(*out, file.data(), file.size());
    *out_n = file.size();
    return 0;
}

extern "C" int sfq_emul_decompress(const uint8_t *sfq_in, size_t n, uint8_t **out, size_t *out_n, uint32_t *status_out) {
    const std::vector<uint8_t> sfq_copy = padded(sfq_in, n);
    const uint8_t *sfq = sfq_copy.data();
    *status_out = 0;
    if (!sfq_is_chunked_container(sfq, n)) { *status_out = SFQ_E_CORRUPT; return 1; }
    SfqFileHeader h; memcpy(&h, sfq, sizeof h);
    if (sfq_index_check(&h, n)) { *status_out = SFQ_E_CORRUPT; return 1; }          // the library's own framing checks (sfq_container.h)
    std::vector<uint8_t> text;
    for (uint64_t c = 0; c < h.nchunks; c++) {
        uint64_t off; memcpy(&off, sfq + h.index_off + 8 * c, 8);
        if (off > n || n - off < sizeof(SfqBlobHeader)) { *status_out = SFQ_E_CORRUPT; return 1; }
        SfqBlobHeader b; memcpy(&b, sfq + off, sizeof b);
        if (sfq_blob_check(&b, off, n)) { *status_out = SFQ_E_CORRUPT; return 1; }
        // (the library is bounded by the caller's output buffer and by device memory; this harness by a fixed limit)
        if (b.out_len > (1ull << 28) || b.nrec > (1u << 24) || b.nbases > (1u << 28) || b.nquals > (1u << 28) || b.hdr_bytes > (1u << 28) ||
            b.big_bases > (1u << 28) || b.big_quals > (1u << 28) || b.big_hdr > (1u << 28)) { *status_out = SFQ_E_CORRUPT; return 1; }
        SfqChunkMeta m; memset(&m, 0, sizeof m);
        m.nrec = b.nrec; m.nbases = b.nbases; m.nquals = b.nquals; m.hdr_bytes = b.hdr_bytes; m.llen = b.llen;
        m.solid = b.solid; m.two_id = b.two_id; m.n_byte = b.n_byte; m.text_len = b.text_len; m.out_len = b.out_len;
        m.nbig = b.nbig; m.big_bases = b.big_bases; m.big_quals = b.big_quals; m.big_hdr = b.big_hdr; m.pad = b.pad;
        const int level = (int)b.level;
        uint64_t soff[SFQ_NSTREAMS]; uint32_t ssize[SFQ_NSTREAMS];
        uint64_t o = off + sizeof b + b.rec_first_len;
        for (int k = 0; k < SFQ_NSTREAMS; k++) { soff[k] = o; ssize[k] = b.ssize[k]; o += b.ssize[k]; }
        uint32_t hbits = sfq_gen_nbuckets(level, b.g_used ? b.g_used : m.nbases, 0);
        void *gt = calloc(1, sfq_gbuckets_bytes(level, hbits));
        uint32_t *qt = (uint32_t *)calloc(1, sfq_qtable_bytes(level));
        uint32_t *pw = (uint32_t *)calloc(1, sfq_pwpool_bytes());
        if (!gt || !qt || !pw) { free(gt); free(qt); free(pw); *status_out = SFQ_E_CORRUPT; return 1; }   // (a wild table hint: the library's cudaMalloc fails the same way)
        static uint32_t lut[SFQ_B2_LUT];
        sfq_b2_lut_fill(lut, 0, 1);
        std::vector<uint32_t> llen(m.nrec), qlen(m.nrec), hlen(m.nrec);
        std::vector<uint8_t> pfg(m.nrec), pfq(m.nrec);
        std::vector<uint64_t> boff(m.nrec), qoff(m.nrec), hoff(m.nrec);
        std::vector<uint8_t> bases((size_t)b.nbases + b.big_bases + 1), quals((size_t)b.nquals + b.big_quals + 1), hdrs((size_t)SFQ_HDR_PLANE(&m));
        sfq_usr_decode_chunk(sfq, ssize, soff, &m, pw, llen.data(), qlen.data(), pfg.data(), pfq.data(), hlen.data(), hoff.data(), boff.data(), qoff.data(),
                             hdrs.data(), hdrs.size(), bases.data(), bases.size() - 1, quals.data(), quals.size() - 1);
        if (m.status) { *status_out = m.status; return 1; }
        uint64_t nb = m.big_bases, nq = m.big_quals;
        for (uint32_t r = 0; r < m.nrec; r++) { if (llen[r] & SFQ_BIG_BIT) continue; boff[r] = nb; qoff[r] = nq; nb += llen[r]; nq += qlen[r]; }
        sfq_gen_decode_chunk(sfq, ssize, soff, &m, level, gt, hbits, pw, llen.data(), boff.data(), bases.data(), lut, SfqStage());
        sfq_gen_apply_exceptions(sfq, ssize, soff, &m, pw, bases.data() + m.big_bases);
        sfq_qlt_decode_chunk(sfq, ssize, soff, &m, level, qt, pw, qlen.data(), qoff.data(), quals.data());
        sfq_rec_decode_chunk(sfq, ssize, soff, &m, pw, sfq + off + sizeof b, b.rec_first_len, hdrs.data(),
                             hdrs.size(), hlen.data(), hoff.data());
        free(gt); free(qt); free(pw);
        if (m.status) { *status_out = m.status; return 1; }
        const uint8_t nbyte = m.n_byte ? m.n_byte : 'N';
        size_t before = text.size();
        for (uint32_t r = 0; r < m.nrec; r++) {       // UsrLoad::save, usrs.cpp:512-529 (the assemble kernel)
            if (hlen[r] & SFQ_BIG_BIT) {              // oversized: verbatim
                const uint32_t hb = hlen[r] & ~SFQ_BIG_BIT;
                const uint8_t *h = hdrs.data() + hoff[r];
                uint32_t cut = 0;
                while (cut < hb && h[cut] != '\n') cut++;
                text.push_back('@'); put_bytes(text, h, cut); text.push_back('\n');
                put_bytes(text, bases.data() + boff[r], llen[r] & ~SFQ_BIG_BIT); text.push_back('\n');
                put_bytes(text, h + cut + 1, hb - cut - 1); text.push_back('\n');
                put_bytes(text, quals.data() + qoff[r], qlen[r] & ~SFQ_BIG_BIT); text.push_back('\n');
                continue;
            }
            text.push_back('@'); put_bytes(text, hdrs.data() + hoff[r], hlen[r]); text.push_back('\n');
            if (m.solid) text.push_back(pfg[r]);
            for (uint32_t i = 0; i < llen[r]; i++) {
                uint8_t cch = bases[boff[r] + i];
                uint8_t q = i < qlen[r] ? quals[qoff[r] + i] : 40;
                if (cch & 0x80) cch &= 0x7f; else if (q == '!') cch = nbyte;
                text.push_back(cch);
            }
            text.push_back('\n'); text.push_back('+');
            if (m.two_id) put_bytes(text, hdrs.data() + hoff[r], hlen[r]);
            text.push_back('\n');
            if (m.solid) text.push_back(pfq[r]);
            put_bytes(text, quals.data() + qoff[r], qlen[r]);
            text.push_back('\n');
        }
        (void)before;
    }
    *out = (uint8_t *)malloc(text.size() + 1);
    memcpy(*out, text.data(), text.size());
    *out_n = text.size();
    return 0;
}
extern "C" void sfq_emul_free(void *p) { free(p); }

// Diagnostics: how many quality contexts / base contexts one chunk touches (sizing of the sparse tables).
extern "C" int sfq_emul_context_counts(const uint8_t *text_in, size_t n, int level, uint64_t *qctx, uint64_t *gctx, uint64_t *nbases) {
    const std::vector<uint8_t> text_copy = padded(text_in, n);
    const uint8_t *text = text_copy.data();
    std::vector<uint64_t> ls;
    ls.push_back(0);
    for (size_t i = 0; i < n; i++) if (text[i] == '\n') ls.push_back(i + 1);
    if ((ls.size() - 1) % 4) return 1;
    SfqChunkMeta m;
    sfq_plan_chunk(text, ls.data(), 0, (ls.size() - 1) / 4, &m);
    if (m.status) return 1;
    SfqArena ar; uint64_t end;
    sfq_arena_layout(&m, 2, 0, &ar, &end);
    std::vector<uint8_t> arena(end);
    uint32_t hbits = sfq_gen_hbits(level, m.nbases, 0);
    uint64_t *gt = (uint64_t *)calloc(1, sfq_gtable_bytes(level, hbits));
    uint32_t *qt = (uint32_t *)calloc(1, sfq_qtable_bytes(level));
    uint32_t *pw = (uint32_t *)calloc(1, sfq_pwpool_bytes());
    sfq_gen_encode_chunk(text, ls.data(), &m, level, gt, hbits, pw, arena.data(), &ar);
    sfq_qlt_encode_chunk(text, ls.data(), &m, level, qt, pw, arena.data(), &ar);
    uint64_t q = 0, g = 0;
    const uint64_t nctx = sfq_qtable_bytes(level) / 256;
    for (uint64_t c = 0; c < nctx; c++) { bool used = false; for (int k = 0; k < 64 && !used; k++) used = qt[c * 64 + k] != 0; q += used; }
    if (level > 1) for (uint64_t k = 0; k < (1ull << hbits); k++) g += gt[k] != 0;
    *qctx = q; *gctx = g; *nbases = m.nbases;
    free(gt); free(qt); free(pw);
    return 0;
}
