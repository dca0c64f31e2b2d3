// host_parsers.cpp — host-side mirrors of the reference's flag parsers and the
// adaptive scaffold clamp.  Pure scalar host code (the reference's is too).
//
//   swg_parse_filter_mode_cli  <- src/main.rs:244-293
//   swg_parse_filter_mode_lib  <- src/library_api.rs:31-63
//   swg_parse_scoring          <- src/main.rs:3485-3492
//   swg_parse_metric_number    <- src/cli.rs:26-61
//   swg_parse_identity_value   <- src/cli.rs:76-130
//   swg_round_nice / swg_clamp_scaffold_params <- src/pansn.rs:176-191, 207-225
//   swg_config_default         <- src/cli.rs:204-276 (clap defaults)
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sweepga_b200.h"
#include "host_util.h"

namespace swg {

// str::parse::<u64>/<usize>: optional '+', then one or more ASCII digits, no overflow.
bool rust_parse_u64(const char *s, size_t len, uint64_t *out) {
    size_t i = 0;
    if (len == 0) return false;
    if (s[0] == '+') i = 1;
    if (i >= len) return false;
    uint64_t v = 0;
    for (; i < len; i++) {
        unsigned d = (unsigned)(s[i] - '0');
        if (d > 9) return false;
        if (v > (~(uint64_t)0 - d) / 10) return false;
        v = v * 10 + d;
    }
    *out = v;
    return true;
}

// str::parse::<f64>: [+-] ( inf | infinity | nan | digits [. digits] [e[+-]digits] | . digits ... ), nothing else.
bool rust_parse_f64(const char *s, size_t len, double *out) {
    if (len == 0 || len > 4000) return false;
    size_t i = 0;
    if (s[i] == '+' || s[i] == '-') i++;
    if (i >= len) return false;
    std::string rest(s + i, len - i), lower = rest;
    for (auto &c : lower) c = (char)tolower((unsigned char)c);
    bool special = (lower == "inf" || lower == "infinity" || lower == "nan");
    if (!special) {
        size_t k = 0, nd = 0;
        while (k < rest.size() && isdigit((unsigned char)rest[k])) { k++; nd++; }
        if (k < rest.size() && rest[k] == '.') {
            k++;
            while (k < rest.size() && isdigit((unsigned char)rest[k])) { k++; nd++; }
        }
        if (nd == 0) return false;
        if (k < rest.size() && (rest[k] == 'e' || rest[k] == 'E')) {
            k++;
            if (k < rest.size() && (rest[k] == '+' || rest[k] == '-')) k++;
            size_t ne = 0;
            while (k < rest.size() && isdigit((unsigned char)rest[k])) { k++; ne++; }
            if (ne == 0) return false;
        }
        if (k != rest.size()) return false;
    }
    std::string z(s, len);
    *out = strtod(z.c_str(), nullptr);
    return true;
}

static std::string to_lower_ascii(const std::string &s) {
    std::string o = s;
    for (auto &c : o) c = (char)tolower((unsigned char)c); // non-ASCII bytes (the UTF-8 of "∞") are untouched
    return o;
}
static std::vector<std::string> split(const std::string &s, char d) {
    std::vector<std::string> out;
    size_t a = 0;
    while (true) {
        size_t b = s.find(d, a);
        if (b == std::string::npos) { out.push_back(s.substr(a)); break; }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    return out;
}
static const char *INF = "\xE2\x88\x9E"; // "∞"

} // namespace swg

using namespace swg;

extern "C" {

void swg_config_default(swg_config *c) {
    std::memset(c, 0, sizeof *c);
    c->min_block_length = 0;
    c->mapping_filter_mode = SWG_MANY_TO_MANY;
    c->mapping_max_per_query = SWG_NO_LIMIT;
    c->mapping_max_per_target = SWG_NO_LIMIT;
    c->scaffold_filter_mode = SWG_MANY_TO_MANY;
    c->scaffold_max_per_query = SWG_NO_LIMIT;
    c->scaffold_max_per_target = SWG_NO_LIMIT;
    c->overlap_threshold = 0.95;
    c->scaffold_overlap_threshold = 0.5;
    c->scaffold_gap = 50000;
    c->min_scaffold_length = 10000;
    c->scaffold_max_deviation = 0;
    c->scoring_function = SWG_SCORE_LOG_LENGTH_IDENTITY;
    c->min_identity = 0.0;
    c->min_scaffold_identity = 0.0;
    c->keep_self = 0;
    c->scaffolds_only = 0;
}

int swg_parse_filter_mode_cli(const char *s, uint8_t *mode, uint64_t *pq, uint64_t *pt) {
    if (!s || !mode || !pq || !pt) return SWG_ERR_ARG;
    std::string orig(s), lower = to_lower_ascii(orig);
    std::string inf(INF);
    auto set = [&](int m, uint64_t q, uint64_t t) { *mode = (uint8_t)m; *pq = q; *pt = t; return SWG_OK; };
    if (lower == "1:1") return set(SWG_ONE_TO_ONE, 1, 1);
    if (lower == "1" || lower == "1:" + inf || lower == "1:infinity" || lower == "1:many") return set(SWG_ONE_TO_MANY, 1, SWG_NO_LIMIT);
    if (lower == inf + ":1" || lower == "infinity:1" || lower == "many:1") return set(SWG_MANY_TO_MANY, SWG_NO_LIMIT, 1);
    if (lower == "many:many" || lower == inf + ":" + inf || lower == "infinity:infinity" || lower == "many" || lower == inf ||
        lower == "infinity" || lower == "-1" || lower == "-1:-1")
        return set(SWG_MANY_TO_MANY, SWG_NO_LIMIT, SWG_NO_LIMIT);
    if (lower.find(':') != std::string::npos) {
        std::vector<std::string> parts = split(lower, ':');
        if (parts.size() == 2) {
            auto side = [&](const std::string &p) -> uint64_t {
                if (p == inf || p == "infinity" || p == "many" || p == "-1") return SWG_NO_LIMIT;
                uint64_t v;
                if (rust_parse_u64(p.data(), p.size(), &v) && v > 0) return v; // 0 rejected -> None
                return SWG_NO_LIMIT;
            };
            uint64_t q = side(parts[0]), t = side(parts[1]);
            int m = (q == 1 && t == 1) ? SWG_ONE_TO_ONE : (q == 1 && t == SWG_NO_LIMIT) ? SWG_ONE_TO_MANY : SWG_MANY_TO_MANY;
            return set(m, q, t);
        }
        return set(SWG_ONE_TO_ONE, 1, 1);
    }
    uint64_t n;
    if (rust_parse_u64(orig.data(), orig.size(), &n)) {
        if (n == 0) return SWG_ERR_PARSE; // the reference calls std::process::exit(1)
        return set(SWG_ONE_TO_MANY, n, SWG_NO_LIMIT);
    }
    return set(SWG_ONE_TO_ONE, 1, 1);
}

int swg_parse_filter_mode_lib(const char *s, uint8_t *mode, uint64_t *pq, uint64_t *pt) {
    if (!s || !mode || !pq || !pt) return SWG_ERR_ARG;
    std::string orig(s), lower = to_lower_ascii(orig);
    auto set = [&](int m, uint64_t q, uint64_t t) { *mode = (uint8_t)m; *pq = q; *pt = t; return SWG_OK; };
    if (lower == "many:many" || lower == "n:n") return set(SWG_MANY_TO_MANY, SWG_NO_LIMIT, SWG_NO_LIMIT);
    std::vector<std::string> parts = split(orig, ':'); // NOT lower-cased in the reference
    if (parts.size() != 2) return set(SWG_ONE_TO_ONE, 1, 1);
    auto side = [&](const std::string &p) -> uint64_t {
        if (p == "many" || p == "n") return SWG_NO_LIMIT;
        uint64_t v;
        if (rust_parse_u64(p.data(), p.size(), &v)) return v; // Some(0) is accepted here
        return SWG_NO_LIMIT;
    };
    uint64_t q = side(parts[0]), t = side(parts[1]);
    if (q == 1 && t == 1) return set(SWG_ONE_TO_ONE, 1, 1);
    if (q == 1) return set(SWG_ONE_TO_MANY, 1, t);
    if (t == 1) return set(SWG_ONE_TO_MANY, q, 1);
    return set(SWG_MANY_TO_MANY, q, t);
}

int swg_parse_scoring(const char *s, uint8_t *scoring) {
    if (!s || !scoring) return SWG_ERR_ARG;
    std::string v(s);
    if (v == "ani" || v == "identity") *scoring = SWG_SCORE_IDENTITY;
    else if (v == "length") *scoring = SWG_SCORE_LENGTH;
    else if (v == "length-ani" || v == "length-identity") *scoring = SWG_SCORE_LENGTH_IDENTITY;
    else if (v == "matches") *scoring = SWG_SCORE_MATCHES;
    else *scoring = SWG_SCORE_LOG_LENGTH_IDENTITY; // incl. "log-length-ani" and anything unknown
    return SWG_OK;
}

int swg_parse_metric_number(const char *s, uint64_t *out) {
    if (!s || !out) return SWG_ERR_ARG;
    size_t len = strlen(s);
    if (len == 0) return SWG_ERR_PARSE;
    char last = s[len - 1];
    bool has_suffix = isalpha((unsigned char)last) && (unsigned char)last < 128;
    size_t nlen = has_suffix ? len - 1 : len;
    double base;
    if (!rust_parse_f64(s, nlen, &base)) return SWG_ERR_PARSE;
    double mult = 1.0;
    if (has_suffix) {
        switch (last) {
        case 'k': case 'K': mult = 1e3; break;
        case 'm': case 'M': mult = 1e6; break;
        case 'g': case 'G': mult = 1e9; break;
        default: return SWG_ERR_PARSE;
        }
    }
    double r = base * mult;
    if (r > 18446744073709551615.0) return SWG_ERR_PARSE;
    // `as u64`: saturating, NaN -> 0
    if (!(r > 0.0)) *out = 0;
    else if (r >= 18446744073709551616.0) *out = ~(uint64_t)0;
    else *out = (uint64_t)r;
    return SWG_OK;
}

int swg_parse_identity_value(const char *s, int has_ani, double ani, double *out) {
    if (!s || !out) return SWG_ERR_ARG;
    std::string value(s), lower = to_lower_ascii(value);
    if (lower.compare(0, 3, "ani") == 0) {
        if (!has_ani) return SWG_ERR_PARSE;
        std::string rem = lower.substr(3);
        if (rem.empty()) { *out = ani; return SWG_OK; }
        size_t p = rem.find('+');
        char sign = 0;
        std::string off;
        if (p != std::string::npos) { sign = '+'; off = rem.substr(p + 1); }
        else if ((p = rem.find('-')) != std::string::npos) { sign = '-'; off = rem.substr(p + 1); }
        if (!sign) { *out = ani; return SWG_OK; }
        double o;
        if (!rust_parse_f64(off.data(), off.size(), &o)) return SWG_ERR_PARSE;
        *out = sign == '+' ? std::fmin(ani + o / 100.0, 1.0) : std::fmax(ani - o / 100.0, 0.0);
        return SWG_OK;
    }
    double v;
    if (!rust_parse_f64(value.data(), value.size(), &v)) return SWG_ERR_PARSE;
    *out = v > 1.0 ? v / 100.0 : v;
    return SWG_OK;
}

// parse_ani_method, src/main.rs:296-331.  SWG_ERR_PARSE <=> None (the caller then uses n50-identity, main.rs:3579).
int swg_parse_ani_method(const char *s, int *method, double *percentile, int *sort) {
    if (!s || !method || !percentile || !sort) return SWG_ERR_ARG;
    const std::string lower = to_lower_ascii(s);
    *percentile = 0.0;
    *sort = SWG_NSORT_IDENTITY;
    if (lower == "all") { *method = SWG_ANI_ALL; return SWG_OK; }
    if (lower == "orthogonal" || lower == "1:1") { *method = SWG_ANI_ORTHOGONAL; return SWG_OK; }
    if (lower.empty() || lower[0] != 'n') return SWG_ERR_PARSE;
    std::vector<std::string> parts = split(lower.substr(1), '-');
    double p;
    if (parts.empty() || !rust_parse_f64(parts[0].data(), parts[0].size(), &p)) return SWG_ERR_PARSE;
    if (!(p > 0.0 && p <= 100.0)) return SWG_ERR_PARSE;
    if (parts.size() > 1) {
        if (parts[1] == "length") *sort = SWG_NSORT_LENGTH;
        else if (parts[1] == "identity") *sort = SWG_NSORT_IDENTITY;
        else if (parts[1] == "score") *sort = SWG_NSORT_SCORE;
        else return SWG_ERR_PARSE;
    }
    *method = SWG_ANI_NPERCENTILE;
    *percentile = p;
    return SWG_OK;
}

uint64_t swg_round_nice(uint64_t v) {
    if (v == 0) return 0;
    uint64_t step = v <= 500 ? 50 : v <= 1000 ? 100 : v <= 3000 ? 200 : 500;
    uint64_t r = (v + step / 2) / step * step;
    return r > step ? r : step;
}

void swg_clamp_scaffold_params(uint64_t user_jump, uint64_t user_mass, int has_avg, uint64_t avg, int adaptive,
                               uint64_t *jump_out, uint64_t *mass_out) {
    *jump_out = user_jump;
    *mass_out = user_mass;
    if (!adaptive || !has_avg || avg == 0) return;
    auto sat_mul = [](uint64_t a, uint64_t b) { return (a != 0 && b > ~(uint64_t)0 / a) ? ~(uint64_t)0 : a * b; };
    uint64_t j10 = sat_mul(avg, 10);
    *jump_out = user_jump < j10 ? user_jump : j10;
    uint64_t m35 = sat_mul(avg, 3) / 5;
    *mass_out = swg_round_nice(user_mass < m35 ? user_mass : m35);
}

const char *swg_version(void) { return "sweepga_b200 0.1.0 (sm_100a)"; }

} // extern "C"
