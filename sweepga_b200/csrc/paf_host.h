// paf_host.h — host-side PAF table shared by the host front end (paf_io.cpp) and the device front end
// (paf_device.cuh, compiled into filter_pipeline.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>

struct swg_paf {
    // the input text (mmap'd or read) — kept so that the writer does not re-read the file
    const char *text = nullptr;
    size_t text_len = 0;
    bool mapped = false;
    int fd = -1;                     // kept open for plain (mmap'd) inputs: the device front end preads from it
    std::vector<char> owned;
    uint64_t n_lines = 0;
    // per record
    std::vector<uint64_t> rank;      // line number
    std::vector<uint64_t> line_off;  // offset of the line in text
    std::vector<uint32_t> line_len;  // length without the line terminator
    std::vector<uint32_t> qid, tid, qs, qe, ts, te, blen, matches;
    std::vector<double> identity;
    std::vector<uint8_t> strand;
    // sequences
    std::vector<std::string> names;
    std::vector<uint32_t> P, P2;
    ~swg_paf() {
        if (mapped && text) munmap((void *)text, text_len);
        if (fd >= 0) close(fd);
    }
};

namespace swg {

// One PAF line the way PafFilter::extract_metadata reads it (src/paf_filter.rs:292-376).
struct PafLine {
    enum Kind { SKIP = 0, OK = 1, ERR_RANGE = 2, ERR_ORDER = 3 };
    const char *qname = nullptr, *tname = nullptr;
    size_t qname_len = 0, tname_len = 0;
    uint32_t qs = 0, qe = 0, ts = 0, te = 0, blen = 0, matches = 0;
    double identity = 0.0;
    uint8_t strand = '+';
};
// `len` excludes the line terminator ("\n" or "\r\n").  SKIP: fewer than 11 fields.
PafLine::Kind paf_parse_line(const char *line, size_t len, PafLine *out);

// One PAF line the way calculate_ani_stats reads it (src/main.rs:412-460, :538-604): f64 matches / block length,
// the FIRST dv:f: tag that parses overrides the matches.  false: the line is skipped ('#', empty, < 11 fields).
struct AniLine {
    const char *qname = nullptr, *tname = nullptr;
    size_t qname_len = 0, tname_len = 0;
    double matches = 0.0, block = 1.0;
    uint64_t qlen = 0, tlen = 0;
};
bool paf_ani_line(const char *line, size_t len, AniLine *out);

// One PAF line the way apply_tree_filter_to_paf reads it (src/tree_filter.rs:219-245): u64 matches / block length
// with unwrap_or(0 / 1).  false: skipped ('#', empty, < 11 fields).  Names come back in `out`.
bool paf_tree_line(const char *line, size_t len, AniLine *out, uint64_t *matches, uint64_t *block);

// SipHash-1-3 with a zero key — std's DefaultHasher::new(), which select_tree_pairs uses for its pseudo-random pairs.
uint64_t siphash13_zero_key(const uint8_t *msg, size_t len);

// select_tree_pairs (src/tree_filter.rs:84-164) over pairs (lo[i], hi[i]) of indices into `genomes`: every genome keeps
// its k nearest and k farthest neighbours by identity, plus the pairs whose hash is below random_fraction * 2^64.
// Ties in identity go by neighbour name (the reference leaves them to HashMap order).  *has_nan: the reference panics.
std::vector<uint8_t> tree_select_pairs(const std::vector<std::string> &genomes, const std::vector<uint32_t> &lo,
                                       const std::vector<uint32_t> &hi, const std::vector<double> &identity, uint64_t k_nearest,
                                       uint64_t k_farthest, double random_fraction, bool *has_nan);

// Opens `path` (plain, .gz or .bgz) and fills text / text_len (mmap when possible); false + message on failure.
bool paf_open_text(const char *path, swg_paf *p, std::string *err);

std::string paf_prefix_P(const std::string &name);  // src/paf_filter.rs:1022-1030
std::string paf_prefix_P2(const std::string &name); // src/plane_sweep_scaffold.rs:13-22

// names (first-appearance order) -> dense ids of P(name) and P2(name), each in first-appearance order
void paf_prefix_ids(const std::vector<std::string> &names, std::vector<uint32_t> *P, std::vector<uint32_t> *P2);

} // namespace swg
