// host_util.h — small host helpers shared by the host-side translation units.
#pragma once
#include <cstddef>
#include <cstdint>

namespace swg {
bool rust_parse_u64(const char *s, size_t len, uint64_t *out); // str::parse::<u64>
bool rust_parse_f64(const char *s, size_t len, double *out);   // str::parse::<f64>
}
