// pack.cuh -- stage many small host arrays through ONE pinned buffer and ONE H2D copy
// (the reference-facing entry points take host pointers; per-array cudaMemcpy calls would be
// launch-latency bound at these sizes).
#pragma once
#include "common.cuh"

namespace plp {

struct Packer {
    struct Item {
        const void *src;
        size_t bytes;
        size_t off;
    };
    std::vector<Item> items;
    size_t total = 0;
    static constexpr size_t kNone = (size_t)-1;

    // returns the byte offset of this array inside the packed buffer (kNone for NULL arrays)
    size_t add(const void *src, size_t bytes) {
        if (src == nullptr) return kNone;
        size_t off = total;
        items.push_back({src, bytes, off});
        total += (bytes + 255) & ~(size_t)255;
        return off;
    }
    // reserve output / scratch space (not copied from the host and NOT cleared: it holds whatever the scratch slot held;
    // every kernel writes its outputs before anything reads them)
    size_t reserve(size_t bytes) {
        size_t off = total;
        total += (bytes + 255) & ~(size_t)255;
        return off;
    }
    plp_status upload(plp_ctx *ctx, int scratch_slot, uint8_t **dbase) {
        void *d = nullptr, *h = nullptr;
        size_t want = total ? total : 256;
        PLP_TRY(ctx_scratch(ctx, scratch_slot, want, &d));
        PLP_TRY(ctx_pinned(ctx, want, &h));
        size_t hi = 0;
        for (const Item &it : items) {
            memcpy((uint8_t *)h + it.off, it.src, it.bytes);
            if (it.off + it.bytes > hi) hi = it.off + it.bytes;
        }
        if (hi) PLP_CUDA_TRY(cudaMemcpyAsync(d, h, hi, cudaMemcpyHostToDevice, ctx->stream));
        *dbase = (uint8_t *)d;
        return PLP_OK;
    }
    template <typename T>
    static T *at(uint8_t *base, size_t off) {
        return off == kNone ? nullptr : reinterpret_cast<T *>(base + off);
    }
};

}  // namespace plp
