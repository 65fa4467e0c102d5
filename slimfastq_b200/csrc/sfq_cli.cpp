// slimfastq-b200: host CLI with the reference's command-line surface (config.cpp:158-201,239-379;
// main.cpp:51-61) over the C ABI.  Option letters, DWIM positionals, stamp detection, refusal to
// overwrite without -O, stdin/stdout piping and exit status follow the reference; all coding is done
// by libsfq_b200.so on the GPU.
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/sfq_b200.h"
#include "sfq_container.h"

static bool g_encode = true;
static std::string g_orig = "";

static void croak(const char *fmt, ...) __attribute__((noreturn));
static void croak(const char *fmt, ...) {        // config.cpp:54-68
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "slimfastq: %s %s: ", g_encode ? "encoding" : "decoding", g_orig.c_str());
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(1);
}

static void usage() {                            // config.cpp:158-201
    printf("\
Usage: \n\
-u  usr-filename : (default: stdin)\n\
-f comp-filename : required - compressed\n\
-d               : decode (instead of encoding) \n\
-O               : silently overwrite existing files\n\
-l level         : compression level 1 to 4 (default is 3 ) \n\
-1, -2, -3, -4   : alias for -l 1, -l 2, etc \n\
 Where levels are:\n\
 1: smallest context tables, yields the worse compression (still much better than gzip)\n\
 2: resonable compression \n\
 3: best compression <default level> \n\
 4: Compress a little more, but costly \n\
\n\
-v               : version : internal version \n\
-h               : help : this message \n\
-s               : stat : information about a compressed file \n\
-q               : suppress extra stats info that could have been seen by -s \n\
-c bytes         : chunk size (default 1048576); every chunk is coded as a standalone file \n\
-g device        : CUDA device index (default 0) \n\
-R               : write the reference's own file format (one chunk; readable by the original slimfastq) \n\
-m MiB           : stream: code the input in segments of about this many MiB of FASTQ (default 2048; 0 = all at once), \n\
                   so pipes of any length run in bounded host memory \n\
\n\
DWIM (Do what I mean) - Intuitive use of 'slimfastq-b200 A B' : \n\
If A appears to be a fastq file, and:\n\
    B does not exists, or -O option is used: compress A to B \n\
If A appears to be a slimfastq file, and: \n\
    B does not exist, or -O option is used: decompress A to B \n\
    B is omitted: decompress A to stdout \n\
Examples: \n\
%% slimfastq-b200 <file.fastq> <new-file.sfq>   : compress <file.fastq> to <new-file.sfq> \n\
%% slimfastq-b200 -1 <file.fastq> <new-file.sfq>: compress <file.fastq> to <new-file.sfq>, using level 1 \n\
%% slimfastq-b200 <file.sfq>                    : decompress <file.sfq> to stdout \n\
%% slimfastq-b200 <file.sfq> <file.fastq>       : decompress <file.sfq> to <file.fastq>\n\
%% gzip -dc <file.fastq.gz> | slimfastq-b200 -f <file.sfq> : convert from gzip to sfq format\n\
Verification example:\n\
%% md5sum <file.fastq>                           : remember checksum \n\
%% slimfastq-b200 <file.fastq> <new-file.sfq>    : compress \n\
%% slimfastq-b200 <new-file.sfq> | md5sum -      : decompress pipe to md5sum, compare checksums \n\
\n");
    exit(0);
}

// Reads a whole stream into a pinned buffer (grown geometrically; stdin has no size).
static uint8_t *slurp(FILE *f, size_t hint, size_t *n_out) {
    size_t cap = hint ? hint + 1 : (64u << 20), n = 0;
    uint8_t *buf = (uint8_t *)sfq_host_alloc(cap);
    if (!buf) croak("cannot allocate %zu bytes of pinned memory (is a CUDA device present?)", cap);
    for (;;) {
        size_t got = fread(buf + n, 1, cap - n, f);
        n += got;
        if (got == 0) break;
        if (n == cap) {
            size_t ncap = cap * 2;
            uint8_t *nb = (uint8_t *)sfq_host_alloc(ncap);
            if (!nb) croak("cannot allocate %zu bytes of pinned memory", ncap);
            memcpy(nb, buf, n);
            sfq_host_free(buf);
            buf = nb; cap = ncap;
        }
    }
    if (ferror(f)) croak("read error: %s", strerror(errno));
    *n_out = n;
    return buf;
}

static void statistics_dump(const uint8_t *p, size_t n) {      // config.cpp:76-85 + filer.cpp:107-118
    SfqFileHeader fh;
    memcpy(&fh, p, sizeof fh);
    fprintf(stderr, ":::: Info ::::\n");
    fprintf(stderr, "%-16s = %s\n", "whoami", "slimfastq");
    fprintf(stderr, "%-16s = %u\n", "version", fh.version);
    fprintf(stderr, "%-16s = %s\n", "format", "b200.c2 (chunked)");
    fprintf(stderr, "%-16s = %u\n", "config.level", fh.level);
    fprintf(stderr, "%-16s = %llu\n", "orig.size", (unsigned long long)fh.orig_size);
    fprintf(stderr, "%-16s = %llu\n", "comp.size", (unsigned long long)n);
    fprintf(stderr, "%-16s = %llu\n", "chunks", (unsigned long long)fh.nchunks);
    fprintf(stderr, "%-16s = %llu\n", "chunk.bytes", (unsigned long long)fh.chunk_bytes);
    static const char *names[SFQ_NSTREAMS] = {"rec", "gen", "qlt", "gen.Ns", "gen.Nn", "rec.x", "usr.x", "usr.x.q", "usr.pfg", "usr.pfq", "usr.lrec", "usr.lgen", "usr.lqlt"};
    unsigned long long tot[SFQ_NSTREAMS] = {0}, nrec = 0, extra = 0;
    if (fh.index_off <= n && fh.nchunks <= (n - fh.index_off) / 8)
        for (uint64_t c = 0; c < fh.nchunks; c++) {
            uint64_t off;
            memcpy(&off, p + fh.index_off + 8 * c, 8);
            if (off > n || n - off < sizeof(SfqBlobHeader)) break;
            SfqBlobHeader b;
            memcpy(&b, p + off, sizeof b);
            for (int k = 0; k < SFQ_NSTREAMS; k++) tot[k] += b.ssize[k];
            nrec += b.nrec; extra += b.extra_hi;
            if (c == 0) {
                fprintf(stderr, "%-16s = %d\n", "llen", b.llen);
                fprintf(stderr, "%-16s = %d\n", "usr.solid", b.solid);
                fprintf(stderr, "%-16s = %d\n", "usr.2id", b.two_id);
                const size_t room = n - off - sizeof b;
                fprintf(stderr, "%-16s = %.*s\n", "rec.first", (int)(b.rec_first_len < room ? b.rec_first_len : room), (const char *)p + off + sizeof b);
            }
        }
    fprintf(stderr, "%-16s = %llu\n", "num_records", nrec);
    if (extra) fprintf(stderr, "%-16s = %llu\n", "qlt.extra.hi", extra);
    fprintf(stderr, "\n:::: Files stream ::::\n i: name      : bytes (all chunks)\n");
    for (int k = 0; k < SFQ_NSTREAMS; k++)
        if (tot[k]) fprintf(stderr, "%2d: %-10s: %llu\n", k + 1, names[k], tot[k]);
    exit(0);
}


// ---- bounded-memory streaming (usrs.cpp:96-122 pages its input through a 1 MiB buffer; here the unit is
// a part of many chunks cut on the chunk grid: each part is coded by one sfq_compress call, its blobs are
// appended to the output as they are produced and one index is written at the end).
struct StreamOut {                       // the container being assembled on disk
    FILE *f = nullptr;
    std::vector<uint64_t> index;
    uint64_t pos = sizeof(SfqFileHeader), orig = 0, out_size = 0, chunk_bytes = 0;
    uint32_t level = 0;
    void begin(FILE *file) { f = file; SfqFileHeader z; memset(&z, 0, sizeof z); if (fwrite(&z, 1, sizeof z, f) != sizeof z) croak("Error writing output: %s", strerror(errno)); }
    void add(const uint8_t *part, size_t n) {      // a whole container: append its blobs
        SfqFileHeader h;
        memcpy(&h, part, sizeof h);
        if (h.index_off > n || h.nchunks > (n - h.index_off) / 8) croak("internal error: bad part container");
        for (uint64_t c = 0; c < h.nchunks; c++) { uint64_t o; memcpy(&o, part + h.index_off + 8 * c, 8); index.push_back(o - sizeof h + pos); }
        const size_t body = (size_t)(h.index_off - sizeof h);
        if (fwrite(part + sizeof h, 1, body, f) != body) croak("Error writing output: %s", strerror(errno));
        pos += body; orig += h.orig_size; out_size += h.out_size; level = h.level; chunk_bytes = h.chunk_bytes;
    }
    void end() {
        if (index.size() && fwrite(index.data(), 8, index.size(), f) != index.size()) croak("Error writing output: %s", strerror(errno));
        SfqFileHeader h;
        sfq_file_header_init(&h, (int)level, orig, index.size(), chunk_bytes, pos, out_size);
        if (fseek(f, 0L, SEEK_SET) || fwrite(&h, 1, sizeof h, f) != sizeof h || fclose(f)) croak("Error writing output: %s", strerror(errno));
    }
};

int main(int argc, char **argv) {
    std::string usr, fil;
    bool overwrite = false, statistics = false, ref_format = false;
    int level = 3, device = 0;
    unsigned long long chunk = 1ull << 20, seg_mib = 2048;
    if (argc == 1) usage();
    const char *short_opt = "qPsvhdOR1234u:f:l:c:g:m:";
    for (int opt = getopt(argc, argv, short_opt); opt != -1; opt = getopt(argc, argv, short_opt))
        switch (opt) {
        case 'u': usr = optarg; break;
        case 'f': fil = optarg; break;
        case 'l': level = (int)strtoll(optarg, 0, 0); break;
        case '1': case '2': case '3': case '4': level = opt - '0'; break;
        case 'c': chunk = strtoull(optarg, 0, 0); break;
        case 'g': device = atoi(optarg); break;
        case 'm': seg_mib = strtoull(optarg, 0, 0); break;
        case 'd': g_encode = false; break;
        case 'O': overwrite = true; break;
        case 'R': ref_format = true; break;
        case 'P': break;                           // profiling cap: accepted, ignored
        case 'q': break;                           // no log.* extras exist in this container
        case 'v': printf("Version 2.04\nInternal format version=%u\n", SFQ_INTERNAL_VERSION); exit(0);
        case 'h': usage(); break;
        case 's': statistics = true; g_encode = false; break;
        default: croak("Ilagal args: use -h for help");
        }
    while (optind < argc) {                        // DWIM guessing, config.cpp:279-325
        char *file = argv[optind++];
        FILE *fh = fopen(file, "rb");
        if (!fh) {
            if (!g_encode && !usr.length()) usr = file;
            else if (g_encode && !fil.length()) fil = file;
            else {
                fprintf(stderr, "What am I suppose to do with '%s'?\n (please specify explicitly with -f/-u prefix)\n(Note: not an existing file)\n", file);
                exit(1);
            }
            continue;
        }
        char initline[20];
        memset(initline, 0, sizeof initline);
        size_t cnt = fread(initline, 1, 19, fh);
        fclose(fh);
        if (cnt && !fil.length() && 0 == strncmp(initline, SFQ_STAMP, 16)) { fil = file; g_encode = !!usr.length(); }
        else if (cnt && !usr.length() && initline[0] == '@') usr = file;
        else if (!usr.length() && !g_encode && (!cnt || overwrite)) usr = file;
        else if (g_encode && usr.length() && !fil.length() && (!cnt || overwrite)) fil = file;
        else {
            fprintf(stderr, "What am I suppose to do with '%s'?\n (please specify explicitly with -f/-u prefix)\n(Note: file exists!)\n", file);
            exit(1);
        }
    }
    if (!fil.length()) { fprintf(stderr, "Missing essential argument: -f\n"); exit(1); }
    const char *wr_flags = overwrite ? "wb" : "wbx";
    g_orig = usr.length() ? usr : (g_encode ? "<< stdin >>" : "");

    sfq_ctx *ctx = nullptr;
    if (!statistics && sfq_create(&ctx, device)) croak("no usable CUDA device (this build has no CPU path)");

    if (g_encode) {
        FILE *out = fopen(fil.c_str(), wr_flags);
        if (!out) { fprintf(stderr, "Can't write file '%s': %s\n", fil.c_str(), strerror(errno)); exit(1); }
        FILE *in = stdin;
        size_t hint = 0;
        if (usr.length()) {
            in = fopen(usr.c_str(), "rb");
            if (!in) { fprintf(stderr, "Can't read file '%s': %s\n", usr.c_str(), strerror(errno)); exit(1); }
            fseek(in, 0L, SEEK_END); hint = (size_t)ftell(in); fseek(in, 0L, SEEK_SET);
        }
        const size_t seg = (size_t)seg_mib << 20;
        if (ref_format || seg == 0 || (hint && hint <= seg)) {             // all at once
            size_t n = 0;
            uint8_t *buf = slurp(in, hint, &n);
            const uint8_t *res = nullptr;
            size_t rn = 0;
            if (ref_format) chunk = ~0ull >> 1;        // the reference's file holds one set of streams: one chunk
            if (sfq_compress(ctx, buf, n, level, chunk, &res, &rn)) { unlink(fil.c_str()); croak("%s", sfq_last_error(ctx)); }
            std::vector<uint8_t> paged;
            if (ref_format) {
                paged.resize(sfq_export_reference_bound(res, rn));
                size_t pn = 0;
                if (sfq_export_reference(res, rn, usr.c_str(), paged.data(), paged.size(), &pn)) { unlink(fil.c_str()); croak("cannot write the reference file format for this input"); }
                res = paged.data(); rn = pn;
            }
            if (fwrite(res, 1, rn, out) != rn || fclose(out)) croak("Error writing output: %s", strerror(errno));
            sfq_host_free(buf);
        } else {                                                           // stream: bounded host memory
            // Every part begins with the first record at or after a line of the chunk grid and is coded with its
            // phase (sfq_set_chunk_phase), so the appended blobs are those of the one-shot container.
            size_t cap = seg + (64u << 20), have = 0;
            uint8_t *buf = (uint8_t *)sfq_host_alloc(cap);
            if (!buf) croak("cannot allocate %zu bytes of pinned memory", cap);
            StreamOut so;
            so.begin(out);
            bool eof = false;
            unsigned long long records_done = 0;
            uint64_t global_off = 0, phase = 0;
            size_t want = seg;                                             // bytes to have buffered before cutting
            while (!eof || have) {
                while (!eof && have < want) {
                    if (want > cap) {
                        uint8_t *nb = (uint8_t *)sfq_host_alloc(want + (64u << 20));
                        if (!nb) croak("cannot allocate %zu bytes of pinned memory", want + (64u << 20));
                        memcpy(nb, buf, have); sfq_host_free(buf); buf = nb; cap = want + (64u << 20);
                    }
                    const size_t got = fread(buf + have, 1, want - have, in);
                    if (got == 0) { if (ferror(in)) croak("read error: %s", strerror(errno)); eof = true; }
                    have += got;
                }
                if (have == 0) break;
                size_t cut = have;
                uint64_t next_phase = 0;
                if (!eof) {
                    const uint64_t grid = !chunk ? 1ull << 20 : chunk < 4096 ? 4096 : chunk;      // as sfq_compress reads it
                    cut = sfq_stream_cut(buf, have, global_off, grid, &next_phase);
                    if (cut == 0) { want = have + seg; continue; }         // no grid line with a record after it yet: read on
                }
                const uint8_t *res = nullptr;
                size_t rn = 0;
                sfq_set_chunk_phase(ctx, phase);
                if (sfq_compress(ctx, buf, cut, level, chunk, &res, &rn)) { unlink(fil.c_str()); croak("%s (in the part after record %llu)", sfq_last_error(ctx), records_done); }
                sfq_stats st;
                sfq_get_stats(ctx, &st);
                records_done += st.nrecords;
                so.add(res, rn);
                memmove(buf, buf + cut, have - cut);
                have -= cut;
                global_off += cut; phase = next_phase; want = seg;
            }
            if (so.index.empty()) { unlink(fil.c_str()); croak("no records were found"); }
            so.end();
            sfq_host_free(buf);
        }
    } else {
        FILE *in = fopen(fil.c_str(), "rb");
        if (!in) { fprintf(stderr, "Can't read file '%s': %s\n", fil.c_str(), strerror(errno)); exit(1); }
        fseek(in, 0L, SEEK_END); size_t hint = (size_t)ftell(in); fseek(in, 0L, SEEK_SET);
        size_t n = 0;
        uint8_t *buf;
        if (statistics) {                          // -s needs no GPU: plain malloc
            buf = (uint8_t *)malloc(hint + 1);
            n = buf ? fread(buf, 1, hint, in) : 0;
        } else buf = slurp(in, hint, &n);
        std::vector<uint8_t> imported;
        const uint8_t *src = buf;
        if (buf && sfq_is_reference_file(buf, n)) {          // a file written by the original slimfastq: one chunk
            imported.resize(sfq_import_reference_bound(n));
            size_t in_n = 0;
            const int rc = sfq_import_reference(buf, n, imported.data(), imported.size(), &in_n);
            if (rc == SFQ_ERR_UNSUPPORTED) croak("%s: reference-format file without orig.size or with oversized-record streams is not supported", fil.c_str());
            if (rc) croak("%s is not a slimfastq file", fil.c_str());
            src = imported.data(); n = in_n;
        } else if (!buf || !sfq_is_chunked_container(buf, n)) croak("%s is not a slimfastq file", fil.c_str());
        if (statistics) statistics_dump(src, n);
        FILE *out = usr.length() ? fopen(usr.c_str(), wr_flags) : stdout;
        if (!out) { fprintf(stderr, "Can't write file '%s': %s\n", usr.c_str(), strerror(errno)); exit(1); }
        SfqFileHeader fh;
        memcpy(&fh, src, sizeof fh);
        const size_t seg = (size_t)seg_mib << 20;
        const bool index_ok = fh.nchunks && fh.index_off <= n && fh.nchunks <= (n - fh.index_off) / 8;
        if (seg == 0 || fh.out_size <= seg || !index_ok) {                  // all at once
            const uint8_t *res = nullptr;
            size_t rn = 0;
            if (sfq_decompress(ctx, src, n, &res, &rn)) { if (usr.length()) unlink(usr.c_str()); croak("%s", sfq_last_error(ctx)); }
            if (fwrite(res, 1, rn, out) != rn) croak("USR: Error writing output");
        } else {                                                            // groups of chunks: bounded output memory
            std::vector<uint64_t> idx(fh.nchunks + 1);
            memcpy(idx.data(), src + fh.index_off, fh.nchunks * 8);
            idx[fh.nchunks] = fh.index_off;
            std::vector<uint8_t> part;
            for (uint64_t c0 = 0; c0 < fh.nchunks;) {
                uint64_t c1 = c0, bytes = 0;
                while (c1 < fh.nchunks) {
                    if (idx[c1] > n || n - idx[c1] < sizeof(SfqBlobHeader) || idx[c1 + 1] < idx[c1] || idx[c1 + 1] > fh.index_off) croak("corrupt container index");
                    SfqBlobHeader b;
                    memcpy(&b, src + idx[c1], sizeof b);
                    if (c1 > c0 && bytes + b.out_len > seg) break;
                    bytes += b.out_len; c1++;
                }
                const uint64_t body = idx[c1] - idx[c0];
                part.resize(sizeof fh + body + (c1 - c0) * 8);
                SfqFileHeader ph;
                sfq_file_header_init(&ph, (int)fh.level, bytes, c1 - c0, fh.chunk_bytes, sizeof fh + body, bytes);
                memcpy(part.data(), &ph, sizeof ph);
                memcpy(part.data() + sizeof ph, src + idx[c0], body);
                for (uint64_t c = c0; c < c1; c++) { const uint64_t o = idx[c] - idx[c0] + sizeof ph; memcpy(part.data() + sizeof ph + body + 8 * (c - c0), &o, 8); }
                const uint8_t *res = nullptr;
                size_t rn = 0;
                if (sfq_decompress(ctx, part.data(), part.size(), &res, &rn)) { if (usr.length()) unlink(usr.c_str()); croak("%s", sfq_last_error(ctx)); }
                if (fwrite(res, 1, rn, out) != rn) croak("USR: Error writing output");
                c0 = c1;
            }
        }
        if (out != stdout ? fclose(out) : fflush(out)) croak("USR: Error writing output");
        sfq_host_free(buf);
    }
    if (ctx) sfq_destroy(ctx);
    return 0;
}
