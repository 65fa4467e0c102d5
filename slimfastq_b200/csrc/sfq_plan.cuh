// Chunk planning: what UsrSave::determine_record (usrs.cpp:186-267) and the structural checks of
// UsrSave::get_record (usrs.cpp:303-390) establish for a standalone file, computed per chunk from
// the line-start table.  One thread per chunk.
#pragma once
#include "sfq_common.cuh"

// Smallest record index whose first byte is at or after `target` (records = groups of 4 lines).
SFQ_HD uint64_t sfq_first_record_at(const uint64_t *ls, uint64_t nrec_total, uint64_t target) {
    uint64_t lo = 0, hi = nrec_total;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (ls[4 * mid] < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The chunk grid.  Chunk c of a whole file holds the records that START in [c*B, (c+1)*B).  When only a
// part of the file is given to a call (a segment of a pipe, the shard of one GPU), `phase` = how far the
// part's first byte lies past the grid line it belongs to, so that its chunks are the very chunks the
// whole-file call would form: slot c of the part starts at local offset max(0, c*B - phase).
// (The part must begin with the first record at or after a grid line, see sfq_b200.h.)
SFQ_HD uint64_t sfq_slot_target(uint64_t c, uint64_t chunk_bytes, uint64_t phase) {
    const uint64_t t = c * chunk_bytes;
    return t > phase ? t - phase : 0;
}
SFQ_HD uint64_t sfq_slot_count(uint64_t n, uint64_t chunk_bytes, uint64_t phase) { return (n + phase + chunk_bytes - 1) / chunk_bytes; }

// Oversized records (usrs.hpp:34-36): the reference's line scans give up after LIMIT - 1 characters
// (`sanity = LIMIT; while (--sanity and c != '\n')`, usrs.cpp:314,334,367), i.e. an id of 8 191+ characters or a base /
// quality line of 65 535+ characters (after the SOLiD prefix) sends the whole record, verbatim, to usr.lrec / usr.lgen /
// usr.lqlt (get_oversized_record, usrs.cpp:269-301) instead of through the models.
SFQ_HD bool sfq_is_big(uint64_t hdr_chars, uint64_t base_chars, uint64_t qual_chars) {
    return hdr_chars >= SFQ_MAX_ID_LLEN - 1 || base_chars >= SFQ_MAX_GN_LLEN - 1 || qual_chars >= SFQ_MAX_GN_LLEN - 1;
}

// Fills `m` for the chunk made of records [r0, r1).  nrec == 0 chunks are left empty (text_len 0).
// `rec_qoff` / `rec_boff` (may be null): [r - r0] = index of record r's first coded quality / base within the chunk.
SFQ_HDN void sfq_plan_chunk(const uint8_t *text, const uint64_t *ls, uint64_t r0, uint64_t r1, SfqChunkMeta *m, uint32_t *rec_qoff = nullptr,
                            uint32_t *rec_boff = nullptr) {
    m->line0 = 4 * r0;
    m->text_off = ls[4 * r0];
    m->text_len = ls[4 * r1] - ls[4 * r0];
    m->out_len = 0;
    m->nrec = (uint32_t)(r1 - r0);
    m->nbases = m->nquals = m->hdr_bytes = 0;
    m->llen = 0; m->solid = 0; m->two_id = 0; m->n_byte = 0; m->pad = 0;
    m->extra_hi = 0; m->q_used = 0; m->g_used = 0; m->status = SFQ_OK; m->status_arg = 0;
    m->nbig = m->big_bases = m->big_quals = m->big_hdr = 0;
    m->first_coded = 0; m->pad2 = 0;
    if (r1 == r0) return;

    uint32_t status = SFQ_OK, arg = 0;
    // ---- determine_record (usrs.cpp:186-267): the first record whose id and base line it can measure decides llen / SOLiD /
    // 2nd-id for the whole chunk; records before it (id of 8 191+ or base line of 65 536+ characters) are put away as oversized
    for (uint64_t r = r0; r < r1; r++) {
        const uint64_t *l = ls + 4 * r;
        const uint64_t hl = l[1] - l[0] - 1, sl64 = l[2] - l[1] - 1;
        if (hl == 0 || text[l[0]] != '@') { status = SFQ_E_AT; arg = (uint32_t)(r - r0 + 1); break; }
        if (hl - 1 >= SFQ_MAX_ID_LLEN - 1 || sl64 >= SFQ_MAX_GN_LLEN) continue;
        const uint32_t sl = (uint32_t)sl64;
        if (sl == 0) { status = SFQ_E_EMPTYSEQ; break; }
        const uint8_t *seq = text + l[1];
        bool solid = false, d_solid = false;
        for (uint32_t i = 1; i < sl && !d_solid && !solid; i++) {
            switch (seq[i] | 0x20) {
            case '0': case '1': case '2': case '3': solid = true; break;
            case 'a': case 'c': case 'g': case 't': d_solid = true; break;
            default: break;
            }
        }
        const uint8_t *plus = text + l[2];
        const uint64_t pl = l[3] - l[2] - 1;
        if (pl == 0 || plus[0] != '+') { status = SFQ_E_PLUS; arg = (uint32_t)(r - r0 + 1); break; }
        bool two = false;
        for (uint64_t i = 1; i < pl; i++) if (plus[i] != ' ') two = true;
        m->solid = solid;
        m->two_id = two;
        m->llen = (int32_t)sl - (solid ? 1 : 0);
        break;
    }
    const uint32_t solid = m->solid;
    uint64_t nb = 0, nq = 0, nh = 0, gb = 0, gq = 0, gh = 0, nbig = 0, out = 0;
    m->first_coded = m->nrec;
    for (uint64_t r = r0; r < r1 && status == SFQ_OK; r++) {
        const uint64_t *l = ls + 4 * r;
        const uint64_t hl = l[1] - l[0] - 1, sl = l[2] - l[1] - 1, pl = l[3] - l[2] - 1, ql = l[4] - l[3] - 1;
        const uint32_t recno = (uint32_t)(r - r0 + 1);
        if (hl == 0 || text[l[0]] != '@') { status = SFQ_E_AT; arg = recno; break; }
        if (rec_qoff) rec_qoff[r - r0] = (uint32_t)nq;
        if (rec_boff) rec_boff[r - r0] = (uint32_t)nb;
        // (a SOLiD line always gives up its first character, usrs.cpp:324-329; an empty one is caught below)
        if (sfq_is_big(hl - 1, sl >= solid ? sl - solid : 0, ql >= solid ? ql - solid : 0)) {
            if (!sfq_is_big(hl - 1, sl >= solid ? sl - solid : 0, 0)) {
                // only the quality line is too long: get_record has read the '+' line by then (usrs.cpp:346-352)
                if (pl == 0 || text[l[2]] != '+') { status = SFQ_E_PLUS; arg = recno; break; }
                if (pl - 1 >= SFQ_MAX_ID_LLEN - 1) { status = SFQ_E_OVERSIZE; arg = recno; break; }
            }
            nbig++;
            gh += (hl - 1) + 1 + pl; gb += sl; gq += ql;
            out += hl + 1 + sl + 1 + pl + 1 + ql + 1;
            continue;
        }
        if (pl == 0 || text[l[2]] != '+') { status = SFQ_E_PLUS; arg = recno; break; }
        if (solid && (sl == 0 || ql == 0)) { status = SFQ_E_TRUNC; arg = recno; break; }
        if (pl - 1 >= SFQ_MAX_ID_LLEN - 1) { status = SFQ_E_OVERSIZE; arg = recno; break; }     // "wierd second id", usrs.cpp:349-352
        if (m->first_coded == m->nrec) m->first_coded = (uint32_t)(r - r0);
        nh += hl - 1;
        nb += sl - solid;
        nq += ql - solid;
        // UsrLoad::save (usrs.cpp:512-529): '@'hdr\n [pf]bases\n '+'[hdr]\n [pf]quals\n
        out += (hl - 1) + 2 + (sl - solid) + (solid + 1) + 2 + (m->two_id ? hl - 1 : 0) + (ql - solid) + (solid + 1);
    }
    if (status == SFQ_OK && ((nb | nq | nh | gb | gq | gh) >= 0xFFFFFF00ull || nb + gb >= 0xFFFFFF00ull || nq + gq >= 0xFFFFFF00ull || nh + gh >= 0xFFFFFF00ull)) { status = SFQ_E_CHUNKSIZE; arg = 0; }      // SfqBlobHeader keeps 32-bit counts
    m->nbases = (uint32_t)nb; m->nquals = (uint32_t)nq; m->hdr_bytes = (uint32_t)nh;
    m->nbig = (uint32_t)nbig; m->big_bases = (uint32_t)gb; m->big_quals = (uint32_t)gq; m->big_hdr = (uint32_t)gh;
    m->out_len = out;
    m->status = status; m->status_arg = arg;
}
