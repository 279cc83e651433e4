// selftest_sharded.cpp -- ONE host process drives several GPUs through the C ABI (vs_sharded_*, include/vscuda.h): the
// shape a `cuda`-tagged server.Search has, being a single Go process (main.go:31).  Builds a striped index over the
// devices given on the command line (default: two stripes -- on devices 0 and 1 when the box has two GPUs, else both on
// device 0), searches it, and checks the hits against a single-device index over all the rows (vs_search), bit for bit.
// Exit code 0 = all checks passed.  Needs a B200 (no CPU fallback): tests/test_gpu_host_cpp.py runs it.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/vscuda.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != VS_OK) {                                                          \
            std::printf("FAIL: %s -> %d (%s)\n", #call, rc_, vs_last_error());       \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    std::vector<int> devices;
    for (int i = 1; i < argc; i++) devices.push_back(std::atoi(argv[i]));
    CHECK(vs_init(0));
    if (devices.empty()) {
        // two stripes: on two GPUs when there are two
        vs_sharded *probe = nullptr;
        const int two[2] = {0, 1};
        if (vs_sharded_create(two, 2, &probe) == VS_OK) {
            vs_sharded_release(probe);
            devices = {0, 1};
        } else {
            devices = {0, 0};
        }
    }
    const size_t n = 20000, d = 768, C = 32, nq = 24, k = 10, nprobe = 6, rb = 8 + d;
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    auto make = [&](size_t rows) {
        std::vector<float> x(rows * d);
        for (auto &v : x) v = nd(rng);
        std::vector<uint8_t> q(rows * rb);
        vs_ctx *c = nullptr;
        if (vs_ctx_create(&c) != VS_OK) std::exit(2);
        if (vs_quantize_f32(c, x.data(), rows, d, q.data()) != VS_OK) std::exit(2);  // compute/quantization.go:82-91
        vs_ctx_destroy(c);
        return q;
    };
    std::vector<uint8_t> rows = make(n), cent = make(C), queries = make(nq);
    vs_ctx *ctx = nullptr;
    CHECK(vs_ctx_create(&ctx));
    // nearest centroid of every row (compute/cosine.go:70-125) = Embedding.CentroidID
    vs_matrix *mc = nullptr, *md = nullptr;
    CHECK(vs_matrix_create(ctx, cent.data(), C, rb, &mc));
    CHECK(vs_matrix_create(ctx, rows.data(), n, rb, &md));
    std::vector<int64_t> idx(n);
    CHECK(vs_argmax_MxN(ctx, mc, md, nullptr, idx.data()));
    std::vector<uint32_t> lists(n);
    std::vector<uint64_t> doc(n);
    for (size_t i = 0; i < n; i++) {
        lists[i] = (uint32_t)idx[i];
        doc[i] = 1000 + i / 2;  // two embeddings per document: one hit per document must hold across shards
    }
    vs_index *one = nullptr;
    CHECK(vs_index_build_assigned(ctx, rows.data(), n, rb, doc.data(), lists.data(), cent.data(), C, &one));
    std::vector<uint64_t> w_ids(nq * k), g_ids(nq * k);
    std::vector<float> w_sims(nq * k), g_sims(nq * k);
    std::vector<int32_t> w_cnt(nq), g_cnt(nq);
    CHECK(vs_search(ctx, one, queries.data(), nq, nprobe, k, w_ids.data(), w_sims.data(), w_cnt.data()));

    vs_sharded *sh = nullptr;
    CHECK(vs_sharded_create(devices.data(), devices.size(), &sh));
    CHECK(vs_sharded_build_assigned(sh, rows.data(), n, rb, doc.data(), lists.data(), cent.data(), C));
    int fails = 0;
    for (size_t batch : {nq, (size_t)1}) {  // a batch, then single queries
        for (size_t q0 = 0; q0 + batch <= nq && q0 < 3 * batch; q0 += batch) {
            CHECK(vs_sharded_search(sh, queries.data() + q0 * rb, batch, nprobe, k, g_ids.data(), g_sims.data(), g_cnt.data()));
            for (size_t q = 0; q < batch; q++) {
                if (g_cnt[q] != w_cnt[q0 + q]) fails++;
                for (int j = 0; j < g_cnt[q]; j++) {
                    if (g_ids[q * k + j] != w_ids[(q0 + q) * k + j]) fails++;
                    if (std::memcmp(&g_sims[q * k + j], &w_sims[(q0 + q) * k + j], 4) != 0) fails++;
                }
            }
        }
    }
    std::printf("sharded selftest: %zu stripes on devices", devices.size());
    for (int dv : devices) std::printf(" %d", dv);
    std::printf(", %zu rows (%zu + ... per stripe), mismatches vs the single-device index: %d\n", vs_sharded_rows(sh),
                vs_sharded_shard_rows(sh, 0), fails);
    vs_sharded_release(sh);
    vs_index_release(one);
    vs_matrix_release(mc);
    vs_matrix_release(md);
    vs_ctx_destroy(ctx);
    if (fails) return 1;
    std::printf("sharded selftest: ok\n");
    return 0;
}
