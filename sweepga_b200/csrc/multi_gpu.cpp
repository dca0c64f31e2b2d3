// multi_gpu.cpp — several GPUs behind ONE C-ABI call (SURVEY 8b "a ctx may own 1-8 devices", 8e).
//
// The filter's groupings are all closed under the genome-pair unit, so the records are partitioned by unit
// (swg_shard_plan: LPT on unit sizes), every device filters its shard with its own context on its own host thread, and the
// only thing that has to be reconciled afterwards is the chain numbering: the kept chains of one unit are numbered
// consecutively both locally and globally (order O3 of SURVEY Appendix B sorts by the unit's first input index A first),
// so each shard reports one (A, first local chain number) pair per unit (swg_last_chain_units), the runs of all shards
// are sorted by the global index of A, and every record's chain number is shifted by its run's offset.  Results are
// identical to a single-GPU call, chain numbers included (tests/test_multi_gpu.py).
//
// This entry point is for callers that hold the whole table in one process (the reference's situation).  The host-side
// split and merge are memory-bound host work; a deployment that wants every GPU fed at PCIe speed runs one process per GPU
// (bench.py --gpus N) and exchanges only the unit runs.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "sweepga_b200.h"

struct swg_multi {
    std::vector<int> devices;
    std::vector<swg_ctx *> ctx;
    std::string err;
};

extern "C" {

swg_multi *swg_multi_create(const int *devices, int n_devices) {
    if (!devices || n_devices < 1) return nullptr;
    swg_multi *m = new (std::nothrow) swg_multi();
    if (!m) return nullptr;
    for (int i = 0; i < n_devices; i++) {
        swg_ctx *c = swg_create(devices[i]);
        if (!c) { // swg_last_error(NULL) holds the reason
            for (swg_ctx *p : m->ctx) swg_destroy(p);
            delete m;
            return nullptr;
        }
        m->devices.push_back(devices[i]);
        m->ctx.push_back(c);
    }
    return m;
}

void swg_multi_destroy(swg_multi *m) {
    if (!m) return;
    for (swg_ctx *p : m->ctx) swg_destroy(p);
    delete m;
}

const char *swg_multi_last_error(const swg_multi *m) { return m ? m->err.c_str() : swg_last_error(nullptr); }
int swg_multi_device_count(const swg_multi *m) { return m ? (int)m->ctx.size() : 0; }

int swg_multi_filter(swg_multi *m, const swg_config *cfg, const swg_mappings *in, swg_result *out, swg_stats *stats) {
    if (!m || !cfg || !in || !out || (in->n && (!out->status || !out->chain_id))) return SWG_ERR_ARG;
    if (in->n && (!in->query_id || !in->target_id)) { m->err = "swg_multi_filter: 32-bit id columns required"; return SWG_ERR_ARG; }
    const int S = (int)m->ctx.size();
    const uint64_t n = in->n;
    if (stats) std::memset(stats, 0, sizeof *stats);
    if (n == 0) return SWG_OK;
    std::vector<uint32_t> shard_of(n);
    std::vector<uint64_t> sizes(S);
    int rc = swg_shard_plan(in, S, shard_of.data(), sizes.data());
    if (rc != SWG_OK) { m->err = "swg_multi_filter: swg_shard_plan failed"; return rc; }
    // index lists (ascending inside a shard: the shard keeps the relative input order, which is all the filter's orders use)
    std::vector<std::vector<uint32_t>> idx(S);
    for (int s = 0; s < S; s++) idx[s].reserve(sizes[s]);
    for (uint64_t i = 0; i < n; i++) idx[shard_of[i]].push_back((uint32_t)i);

    struct Shard {
        std::vector<uint32_t> qid, tid, qs, qe, ts, te, bl, mt, chain, runA, runK;
        std::vector<double> ident, score;
        std::vector<uint8_t> strand, status;
        swg_stats st;
        int rc = SWG_OK;
        std::string err;
    };
    std::vector<Shard> sh(S);
    auto work = [&](int s) {
        Shard &w = sh[s];
        const std::vector<uint32_t> &ix = idx[s];
        const size_t k = ix.size();
        std::memset(&w.st, 0, sizeof w.st);
        if (k == 0) return;
        auto g32 = [&](const uint32_t *src, std::vector<uint32_t> &dst) { dst.resize(k); for (size_t i = 0; i < k; i++) dst[i] = src[ix[i]]; };
        g32(in->query_id, w.qid); g32(in->target_id, w.tid); g32(in->query_start, w.qs); g32(in->query_end, w.qe);
        g32(in->target_start, w.ts); g32(in->target_end, w.te); g32(in->block_length, w.bl); g32(in->matches, w.mt);
        if (in->identity) { w.ident.resize(k); for (size_t i = 0; i < k; i++) w.ident[i] = in->identity[ix[i]]; }
        if (in->score) { w.score.resize(k); for (size_t i = 0; i < k; i++) w.score[i] = in->score[ix[i]]; }
        w.strand.resize(k);
        for (size_t i = 0; i < k; i++) w.strand[i] = in->strand[ix[i]];
        w.status.assign(k, 0);
        w.chain.assign(k, 0);
        swg_mappings sub = *in;
        sub.n = k;
        sub.query_id = w.qid.data(); sub.target_id = w.tid.data(); sub.query_start = w.qs.data(); sub.query_end = w.qe.data();
        sub.target_start = w.ts.data(); sub.target_end = w.te.data(); sub.block_length = w.bl.data(); sub.matches = w.mt.data();
        sub.identity = in->identity ? w.ident.data() : nullptr;
        sub.score = in->score ? w.score.data() : nullptr;
        sub.strand = w.strand.data();
        sub.query_id16 = sub.target_id16 = nullptr;
        swg_result res{w.status.data(), w.chain.data()};
        w.rc = swg_filter(m->ctx[s], cfg, &sub, &res, &w.st);
        if (w.rc != SWG_OK) { w.err = swg_last_error(m->ctx[s]); return; }
        uint64_t nu = 0;
        w.rc = swg_last_chain_units(m->ctx[s], 0, nullptr, nullptr, &nu);
        if (w.rc != SWG_OK) { w.err = swg_last_error(m->ctx[s]); return; }
        w.runA.resize(nu);
        w.runK.resize(nu);
        if (nu) w.rc = swg_last_chain_units(m->ctx[s], nu, w.runA.data(), w.runK.data(), &nu);
        if (w.rc != SWG_OK) w.err = swg_last_error(m->ctx[s]);
    };
    {
        std::vector<std::thread> th;
        for (int s = 1; s < S; s++) th.emplace_back(work, s);
        work(0);
        for (auto &t : th) t.join();
    }
    for (int s = 0; s < S; s++)
        if (sh[s].rc != SWG_OK) { m->err = "device " + std::to_string(m->devices[s]) + ": " + sh[s].err; return sh[s].rc; }
    // merge the numberings: runs of all shards by the global index of A
    struct Run { uint64_t a; int s; uint32_t r; uint64_t count; };
    std::vector<Run> runs;
    for (int s = 0; s < S; s++) {
        const Shard &w = sh[s];
        for (size_t r = 0; r < w.runA.size(); r++) {
            const uint64_t next = r + 1 < w.runK.size() ? w.runK[r + 1] : w.st.n_chains_kept + 1;
            runs.push_back(Run{idx[s][w.runA[r]], s, (uint32_t)r, next - w.runK[r]});
        }
    }
    std::sort(runs.begin(), runs.end(), [](const Run &x, const Run &y) { return x.a < y.a; });
    std::vector<std::vector<int64_t>> delta(S);
    for (int s = 0; s < S; s++) delta[s].resize(sh[s].runA.size());
    uint64_t excl = 0;
    for (const Run &r : runs) {
        delta[r.s][r.r] = (int64_t)(excl + 1) - (int64_t)sh[r.s].runK[r.r];
        excl += r.count;
    }
    auto scatter = [&](int s) {
        const Shard &w = sh[s];
        const std::vector<uint32_t> &ix = idx[s];
        for (size_t i = 0; i < ix.size(); i++) {
            out->status[ix[i]] = w.status[i];
            uint32_t k = w.chain[i];
            if (k) {
                const size_t r = (size_t)(std::upper_bound(w.runK.begin(), w.runK.end(), k) - w.runK.begin()) - 1;
                k = (uint32_t)((int64_t)k + delta[s][r]);
            }
            out->chain_id[ix[i]] = k;
        }
    };
    {
        std::vector<std::thread> th;
        for (int s = 1; s < S; s++) th.emplace_back(scatter, s);
        scatter(0);
        for (auto &t : th) t.join();
    }
    if (stats) {
        for (int s = 0; s < S; s++) {
            const swg_stats &a = sh[s].st;
            stats->n_input += a.n_input; stats->n_stage1 += a.n_stage1; stats->n_after_sweep += a.n_after_sweep; stats->n_chains += a.n_chains;
            stats->n_chains_after_mass += a.n_chains_after_mass; stats->n_chains_kept += a.n_chains_kept; stats->n_anchors += a.n_anchors;
            stats->n_rescued += a.n_rescued; stats->n_kept += a.n_kept; stats->score_near_ties += a.score_near_ties;
            stats->gpu_launches += a.gpu_launches; stats->exact_rerank |= a.exact_rerank; stats->n_dirty_groups += a.n_dirty_groups;
            stats->h2d_bytes += a.h2d_bytes; stats->d2h_bytes += a.d2h_bytes;
            stats->ms_h2d = std::max(stats->ms_h2d, a.ms_h2d); stats->ms_device = std::max(stats->ms_device, a.ms_device);
            stats->ms_d2h = std::max(stats->ms_d2h, a.ms_d2h);
        }
    }
    return SWG_OK;
}

} // extern "C"
