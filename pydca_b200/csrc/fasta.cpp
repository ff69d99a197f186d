// fasta.cpp -- FASTA ingest, integer encoding and first-occurrence de-duplication.
//
// Replaces  PlmDCA::readSequencesFromFile   pydca/plmdca/plmdca_numerics.cpp:685-767
//           alignment_letter2int            pydca/fasta_reader/fasta_reader.py:122-163
//
// Semantics kept: letters are upper-cased (:752 / fasta_reader.py:141), encoded
// with the C++ tables (protein ACDEFGHIKLMNPQRSTVWY -> 0..19, RNA ACGU -> 0..3,
// everything else -> gap = q-1), and only the first occurrence of each distinct
// sequence is kept, in file order (:756 / fasta_reader.py:153).
// Deliberate deltas ("reader semantics"):
//   * records may span several lines (Biopython semantics); the C++ reader
//     would treat every line as a sequence.  Identical for unwrapped files.
//   * letters missing from the C++ map (e.g. 'T' in RNA) are gaps, as in the
//     Python reader; the C++ reader aborts on them (:752).
//   * de-duplication is hash-then-verify, O(N L) instead of O(N^2 L).
// Large files are cut into pieces in front of header lines; the pieces are translated and hashed on host
// threads, first occurrences are decided in one serial pass over the hashes in file order, and the kept
// records are copied into the output by one thread per piece.
#include "dca_common.cuh"
#include "dca_internal.h"

#include <cctype>
#include <cstdio>
#include <string>
#include <algorithm>
#include <chrono>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <system_error>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace dca {

int encode_residue(int biomolecule, int c)
{
    static const char prot[] = "ACDEFGHIKLMNPQRSTVWY";
    static const char rna[] = "ACGU";
    const char *alpha = (biomolecule == 1) ? prot : rna;
    const int q = (biomolecule == 1) ? 21 : 5;
    const int u = std::toupper(c);
    if (u == 0) return q - 1;
    const char *p = std::strchr(alpha, u);
    return p ? (int)(p - alpha) : q - 1;
}

// 64-bit hash of an encoded record, 8 bytes per step (only used to find candidate duplicates:
// equality is always confirmed with memcmp)
static uint64_t record_hash(const uint8_t *p, size_t n)
{
    uint64_t h = 0x9e3779b97f4a7c15ull ^ (uint64_t)n;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        std::memcpy(&w, p + i, 8);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    uint64_t w = 0;
    if (i < n) std::memcpy(&w, p + i, n - i);
    h = (h ^ w) * 0xc4ceb9fe1a85ec53ull;
    return h ^ (h >> 29);
}

namespace {

// one contiguous piece of the file, starting at the beginning of a line (piece 0: the file's first byte,
// every other piece: a '>' header line), parsed on its own thread
struct Piece {
    const char *begin = nullptr, *end = nullptr;
    std::vector<uint8_t> recs;         // the piece's records in file order, L bytes each (duplicates included)
    std::vector<uint64_t> hashes;
    long lines = 0;                    // lines in the piece
    long bad_line = 0;                 // first malformed record: line (within the piece) where it ended, 0 = none
    int bad_fill = 0;
};

void parse_piece(Piece &pc, const uint8_t *lut, int L)
{
    std::vector<uint8_t> rec((size_t)L);
    int fill = -1;                     // -1: no record open yet
    long line_no = 0;
    pc.recs.reserve((size_t)(pc.end - pc.begin));
    auto flush = [&]() {
        if (fill <= 0) return true;    // empty record: skipped like the reference's empty lines
        if (fill < L) { pc.bad_line = line_no; pc.bad_fill = fill; return false; }
        pc.hashes.push_back(record_hash(rec.data(), (size_t)L));
        pc.recs.insert(pc.recs.end(), rec.begin(), rec.end());
        return true;
    };
    const char *p = pc.begin, *end = pc.end;
    bool ok = true;
    while (p < end && ok) {
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        const char *stop = eol ? eol : end;
        ++line_no;
        if (p < stop && *p == '>') { ok = flush(); fill = 0; }
        else {
            if (fill < 0 && p < stop) fill = 0;     // headerless file: one record
            const size_t len = (size_t)(stop - p);
            // lines without blanks (every line of a normal alignment) are translated in one
            // branch-free run; the byte-by-byte loop below was 60 % of the reader's time
            if (!std::memchr(p, '\r', len) && !std::memchr(p, ' ', len) && !std::memchr(p, '\t', len)) {
                const size_t room = (fill < L) ? (size_t)(L - fill) : 0, take = len < room ? len : room;
                uint8_t *dst = rec.data() + fill;
                const unsigned char *src = reinterpret_cast<const unsigned char *>(p);
                for (size_t k = 0; k < take; ++k) dst[k] = lut[src[k]];
                fill += (int)len;
            } else {
                for (const char *c = p; c < stop; ++c) {
                    const uint8_t v = lut[(unsigned char)*c];
                    if (v == 255) continue;
                    if (fill < L) rec[(size_t)fill] = v;
                    ++fill;
                }
            }
        }
        p = eol ? eol + 1 : end;
    }
    if (ok) flush();
    pc.lines = line_no;
}

}  // namespace

long read_fasta_unique(const char *path, int biomolecule, int L, ByteVec &out)
{
    const auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) { if (getenv("DCA_B200_READER_TRACE")) fprintf(stderr, "[reader] %s %.2f ms\n", w, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count()); };
    // The whole file at once: mapped when it is a regular file (no copy, the parser threads fault
    // the pages in), read in pieces otherwise (pipe).
    std::vector<char> text;
    const char *tb = nullptr;
    size_t tsize = 0;
    void *map = MAP_FAILED;
    size_t map_len = 0;
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { set_error("Unable to open file %s", path); return -1; }
        struct stat sb;
        if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            map_len = (size_t)sb.st_size;
            map = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
        }
        if (map != MAP_FAILED) { tb = static_cast<const char *>(map); tsize = map_len; }
        else {
            char buf[1 << 16];
            ssize_t got;
            while ((got = read(fd, buf, sizeof buf)) > 0) text.insert(text.end(), buf, buf + got);
            tb = text.data(); tsize = text.size();
        }
        close(fd);
    }
    struct Unmap { void *m; size_t n; ~Unmap() { if (m != MAP_FAILED) munmap(m, n); } } unmap{map, map_len};
    lap("read");
    uint8_t lut[256];                              // 255 = white space inside a line (skipped)
    for (int c = 0; c < 256; ++c) lut[c] = (uint8_t)encode_residue(biomolecule, c);
    lut[(unsigned char)'\r'] = lut[(unsigned char)' '] = lut[(unsigned char)'\t'] = 255;

    // Pieces of >= 128 KB cut in front of header lines ("\n>"), one host thread each (translation and
    // hashing are the reader's time: 15 ms of a 31 ms mfDCA call at N = 20 000, L = 300 on one
    // thread); the first-occurrence de-duplication below walks the pieces in file order.
    const char *te = tb + tsize;
    size_t want = tsize / (1u << 17);               // 128 KB per piece and thread (a thread start costs ~50 us)
    const unsigned hc = std::thread::hardware_concurrency();
    want = std::min<size_t>(want, std::min<unsigned>(hc ? hc : 1u, 16u));
    if (const char *e = getenv("DCA_B200_READER_THREADS")) want = (size_t)std::min(256, std::max(1, atoi(e)));
    want = std::max<size_t>(want, 1);
    std::vector<Piece> pieces;
    {
        const char *start = tb;
        for (size_t t = 1; t < want && start < te; ++t) {
            const char *from = tb + tsize * t / want;
            if (from <= start) continue;
            const char *cut = nullptr;
            for (const char *q = from; q + 1 < te; ) {
                q = static_cast<const char *>(std::memchr(q, '\n', (size_t)(te - q - 1)));
                if (!q) break;
                if (q[1] == '>') { cut = q + 1; break; }
                ++q;
            }
            if (!cut) break;
            Piece pc; pc.begin = start; pc.end = cut;
            pieces.push_back(std::move(pc));
            start = cut;
        }
        Piece last; last.begin = start; last.end = te;
        pieces.push_back(std::move(last));
    }
    if (pieces.size() == 1) parse_piece(pieces[0], lut, L);
    else {
        std::vector<std::thread> th;
        size_t started = 1;
        try {
            for (; started < pieces.size(); ++started) th.emplace_back(parse_piece, std::ref(pieces[started]), lut, L);
        } catch (const std::system_error &) {
            // no more threads to be had: this thread parses the rest
        }
        parse_piece(pieces[0], lut, L);
        for (size_t t = started; t < pieces.size(); ++t) parse_piece(pieces[t], lut, L);
        for (auto &x : th) x.join();
    }
    lap("parsed");
    {   // the first malformed record in file order is the one the single pass would have stopped at
        long lines_before = 0;
        for (const Piece &pc : pieces) {
            if (pc.bad_line) {
                set_error("%s: record ending at line %ld has %d residues, expected %d", path, lines_before + pc.bad_line,
                          pc.bad_fill, L);
                return -2;
            }
            lines_before += pc.lines;
        }
    }

    size_t total = 0;
    for (const Piece &pc : pieces) total += pc.hashes.size();
    // First occurrences in file order: a serial pass over the hashes only (open addressing, equality confirmed
    // on the pieces' own bytes) gives every kept record its row; the bytes are then copied by one thread per
    // piece into the output, which is sized without being zero-filled.
    size_t cap = 1024;
    while (cap < total * 2 + 16) cap <<= 1;
    std::vector<const uint8_t *> slot(cap, nullptr);
    std::vector<uint64_t> slot_hash(cap, 0);
    std::vector<std::vector<long>> dest(pieces.size());    // per piece: output row of each record, -1 = duplicate
    long n = 0;
    for (size_t t = 0; t < pieces.size(); ++t) {
        const Piece &pc = pieces[t];
        dest[t].resize(pc.hashes.size());
        for (size_t r = 0; r < pc.hashes.size(); ++r) {
            const uint8_t *rec = pc.recs.data() + r * (size_t)L;
            const uint64_t h = pc.hashes[r];
            size_t k = (size_t)h & (cap - 1);
            bool dup = false;
            while (slot[k]) {                      // first occurrence wins (plmdca_numerics.cpp:756)
                if (slot_hash[k] == h && std::memcmp(slot[k], rec, (size_t)L) == 0) { dup = true; break; }
                k = (k + 1) & (cap - 1);
            }
            if (dup) { dest[t][r] = -1; continue; }
            slot[k] = rec; slot_hash[k] = h;
            dest[t][r] = n++;
        }
    }
    out.clear();
    out.resize((size_t)n * (size_t)L);
    auto copy_piece = [&](size_t t) {
        const Piece &pc = pieces[t];
        const std::vector<long> &d = dest[t];
        for (size_t r = 0; r < d.size(); ++r)
            if (d[r] >= 0) std::memcpy(out.data() + (size_t)d[r] * L, pc.recs.data() + r * (size_t)L, (size_t)L);
    };
    if (pieces.size() == 1) copy_piece(0);
    else {
        std::vector<std::thread> th;
        size_t started = 1;
        try {
            for (; started < pieces.size(); ++started) th.emplace_back(copy_piece, started);
        } catch (const std::system_error &) {
        }
        copy_piece(0);
        for (size_t t = started; t < pieces.size(); ++t) copy_piece(t);
        for (auto &x : th) x.join();
    }
    lap("merged");
    if (n == 0) { set_error("%s: no sequences found", path); return -3; }
    return n;
}

}  // namespace dca
