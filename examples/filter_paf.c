/* The boundary from plain C: parse CLI-style flags with the library's own parsers, filter a PAF file to a tagged PAF file.
 *
 *   gcc -Iinclude examples/filter_paf.c -o filter_paf sweepga_b200/libsweepga_b200.so -Wl,-rpath,$PWD/sweepga_b200
 *   ./filter_paf in.paf out.paf [num_mappings (e.g. 1:1)] [scaffold_filter (e.g. 1:1)] [device]
 *
 * This is the call sequence the patched `PafFilter::filter_paf` (patches/apply_filters.patch) makes through the
 * Rust sys crate; here without Rust.  There is no CPU path: without an sm_100 device swg_create fails and says why.
 * `--plan` instead of a PAF runs only host-side entry points (no GPU needed): used by tests/test_host.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sweepga_b200.h"

static int host_only_plan(void) {
    /* swg_shard_plan_units: LPT of genome-pair unit sizes onto shards — what a multi-GPU driver calls before it loads anything */
    const uint64_t sizes[6] = {50, 10, 40, 30, 20, 60};
    uint32_t shard_of[6];
    uint64_t load[2];
    if (swg_shard_plan_units(6, sizes, 2, shard_of, load) != SWG_OK) return 1;
    printf("version %s\nshard loads %llu %llu\n", swg_version(), (unsigned long long)load[0], (unsigned long long)load[1]);
    swg_config cfg;
    swg_config_default(&cfg);
    printf("default scaffold_gap %llu min_scaffold_length %llu\n", (unsigned long long)cfg.scaffold_gap, (unsigned long long)cfg.min_scaffold_length);
    return load[0] + load[1] == 210 ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "--plan") == 0) return host_only_plan();
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.paf out.paf [num_mappings] [scaffold_filter] [device]\n       %s --plan\n", argv[0], argv[0]);
        return 2;
    }
    swg_config cfg;
    swg_config_default(&cfg);
    if (argc > 3 && swg_parse_filter_mode_cli(argv[3], &cfg.mapping_filter_mode, &cfg.mapping_max_per_query, &cfg.mapping_max_per_target) != SWG_OK) {
        fprintf(stderr, "bad --num-mappings value '%s'\n", argv[3]);
        return 2;
    }
    if (argc > 4 && swg_parse_filter_mode_cli(argv[4], &cfg.scaffold_filter_mode, &cfg.scaffold_max_per_query, &cfg.scaffold_max_per_target) != SWG_OK) {
        fprintf(stderr, "bad --scaffold-filter value '%s'\n", argv[4]);
        return 2;
    }
    swg_ctx *ctx = swg_create(argc > 5 ? atoi(argv[5]) : 0);
    if (!ctx) {
        fprintf(stderr, "swg_create: %s\n", swg_last_error(NULL));
        return 3;
    }
    swg_stats st;
    const int rc = swg_filter_file(ctx, &cfg, argv[1], argv[2], /*keep_self=*/0, &st);
    if (rc != SWG_OK) fprintf(stderr, "swg_filter_file: %d: %s\n", rc, swg_last_error(ctx));
    else
        printf("%llu of %llu records kept, %llu chains, filter %.3f ms on the device (%llu kernel launches)\n", (unsigned long long)st.n_kept,
               (unsigned long long)st.n_input, (unsigned long long)st.n_chains_kept, st.ms_device, (unsigned long long)st.gpu_launches);
    swg_destroy(ctx);
    return rc == SWG_OK ? 0 : 1;
}
