/* nwb_batch_api.inl -- batch entry points (placeholder until implemented). */
struct nwb_batch { int64_t n; };
extern "C" int nwb_fill_batch(const char *, const int64_t *, const char *, const int64_t *, int64_t,
                              int, int, int, unsigned, int, nwb_batch **out)
{
    if (out) *out = nullptr;
    return NWB_ERR_UNSUPPORTED;
}
extern "C" void nwb_batch_free(nwb_batch *b) { delete b; }
extern "C" int64_t nwb_batch_size(const nwb_batch *b) { return b ? b->n : 0; }
extern "C" int32_t nwb_batch_opt_score(const nwb_batch *, int64_t) { return 0; }
extern "C" uint32_t nwb_batch_branch_count(const nwb_batch *, int64_t) { return 0; }
extern "C" uint64_t nwb_batch_count_u64(const nwb_batch *, int64_t) { return 0; }
extern "C" const uint8_t *nwb_batch_arrow_rows(const nwb_batch *, int64_t, size_t *) { return nullptr; }
extern "C" float nwb_batch_kernel_ms(const nwb_batch *) { return 0.f; }
