// neural_stubs.cu -- placeholders for entry points whose kernels are not built into this library
// configuration.  They fail loudly; there is no fallback path.
#include "npd_common.cuh"

#ifndef NPD_HAVE_GRU
NPD_API int npd_gru_create(int, int, const float *, const float *, const float *, const float *,
                           const float *, const float *, const float *, const float *, const float *,
                           const float *, npd_gru_t **)
{
    npd_set_error("npd_gru_create: GRU kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
NPD_API int npd_gru_destroy(npd_gru_t *) { return NPD_OK; }
NPD_API int npd_gru_set_head_mlp(npd_gru_t *, int, int, const float *)
{
    npd_set_error("npd_gru_set_head_mlp: GRU kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
NPD_API int npd_gru_decode_h0(const npd_gru_t *, const npd_code_t *, const float *, const float *, const float *, const float *,
                              float *, float *, int64_t, void *, size_t, void *)
{
    npd_set_error("npd_gru_decode_h0: GRU kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
NPD_API size_t npd_gru_workspace_bytes(const npd_gru_t *, int64_t) { return 0; }
NPD_API int npd_gru_decode(const npd_gru_t *, const npd_code_t *, const float *, const float *, const float *, float *,
                           float *, int64_t, void *, size_t, void *)
{
    npd_set_error("npd_gru_decode: GRU kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
#endif

#ifndef NPD_HAVE_CONV
NPD_API int npd_conv_create(int, int, const float *, size_t, npd_conv_t **)
{
    npd_set_error("npd_conv_create: conv kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
NPD_API int npd_conv_destroy(npd_conv_t *) { return NPD_OK; }
NPD_API size_t npd_conv_workspace_bytes(const npd_conv_t *, int64_t) { return 0; }
NPD_API int npd_conv_decode(const npd_conv_t *, const float *, float *, int64_t, void *, size_t, void *)
{
    npd_set_error("npd_conv_decode: conv kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
NPD_API int npd_conv_forward(const npd_conv_t *, const float *, float *, float *, int64_t, void *, size_t, void *)
{
    npd_set_error("npd_conv_forward: conv kernel not built into this libnpd.so");
    return NPD_EUNSUPPORTED;
}
#endif
