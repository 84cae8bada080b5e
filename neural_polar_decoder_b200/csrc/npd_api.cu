// npd_api.cu -- library plumbing: versioning, errors, device properties, code objects.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "npd_common.cuh"

static thread_local char g_err[512] = "";

void npd_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

NPD_API int npd_version(void) { return NPD_VERSION; }
NPD_API const char *npd_last_error(void) { return g_err; }

int npd_get_device_props(DeviceProps *p)
{
    static std::mutex mu;
    static DeviceProps cache[64];
    static bool have[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        npd_set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return NPD_ECUDA;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && have[dev]) {
        *p = cache[dev];
        return NPD_OK;
    }
    DeviceProps d{};
    d.device = dev;
    if (cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        npd_set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
        return NPD_ECUDA;
    }
    if (dev < 64) {
        cache[dev] = d;
        have[dev] = true;
    }
    *p = d;
    return NPD_OK;
}

NPD_API int npd_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len)
{
    DeviceProps d;
    if (npd_get_device_props(&d)) return NPD_ECUDA;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    if (name && name_len > 0) {
        cudaDeviceProp prop;
        NPD_CHECK_CUDA(cudaGetDeviceProperties(&prop, d.device));
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return NPD_OK;
}

NPD_API int npd_code_create(int n, int K, const int32_t *h_info, float infty, uint32_t pac_g,
                            npd_code_t **out)
{
    NPD_REQUIRE(out, "npd_code_create: null out");
    *out = nullptr;
    NPD_REQUIRE(n >= 1 && n <= 12, "npd_code_create: n=%d outside 1..12", n);
    const int N = 1 << n;
    NPD_REQUIRE(K >= 0 && K <= N, "npd_code_create: K=%d outside 0..%d", K, N);
    NPD_REQUIRE(K == 0 || h_info, "npd_code_create: null info set");
    for (int k = 0; k < K; ++k) {
        NPD_REQUIRE(h_info[k] >= 0 && h_info[k] < N, "npd_code_create: info[%d]=%d out of range", k,
                    h_info[k]);
        NPD_REQUIRE(k == 0 || h_info[k] > h_info[k - 1],
                    "npd_code_create: info positions must be strictly increasing");
    }
    DeviceProps dp;
    if (npd_get_device_props(&dp)) return NPD_ECUDA;

    npd_code *c = (npd_code *)calloc(1, sizeof(npd_code));
    if (!c) return NPD_ENOMEM;
    c->n = n; c->N = N; c->K = K; c->infty = infty; c->pac_g = pac_g;
    c->device = dp.device; c->sm_count = dp.sm_count;
    if (pac_g) {
        int M = 0;
        while ((pac_g >> M) != 0) ++M;  // floor(log2 g) + 1 (pac_code.py:101)
        c->pac_M = M;
        uint32_t taps = 0;
        for (int j = 1; j < M; ++j)  // g_array[j] = bit (M-1-j) of g, MSB first (pac_code.py:102)
            if ((pac_g >> (M - 1 - j)) & 1u) taps |= 1u << (j - 1);
        c->pac_taps = taps;
    }
    const int NW = (N + 31) / 32;
    uint32_t *fw = (uint32_t *)calloc(NW, 4), *iw = (uint32_t *)calloc(NW, 4);
    c->h_info = (int32_t *)malloc(sizeof(int32_t) * (K > 0 ? K : 1));
    if (!fw || !iw || !c->h_info) { free(fw); free(iw); free(c->h_info); free(c); return NPD_ENOMEM; }
    for (int i = 0; i < N; ++i) fw[i >> 5] |= 1u << (i & 31);
    for (int k = 0; k < K; ++k) {
        c->h_info[k] = h_info[k];
        fw[h_info[k] >> 5] &= ~(1u << (h_info[k] & 31));
        iw[h_info[k] >> 5] |= 1u << (h_info[k] & 31);
    }
    cudaError_t e = cudaMalloc(&c->d_info, sizeof(int32_t) * (K > 0 ? K : 1));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_frozen_words, 4 * NW);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_info_words, 4 * NW);
    if (e == cudaSuccess && K > 0)
        e = cudaMemcpy(c->d_info, c->h_info, sizeof(int32_t) * K, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_frozen_words, fw, 4 * NW, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_info_words, iw, 4 * NW, cudaMemcpyHostToDevice);
    free(fw);
    free(iw);
    if (e != cudaSuccess) {
        npd_set_error("npd_code_create: %s", cudaGetErrorString(e));
        npd_code_destroy(c);
        return NPD_ECUDA;
    }
    *out = c;
    return NPD_OK;
}

NPD_API int npd_code_destroy(npd_code_t *c)
{
    if (!c) return NPD_OK;
    if (c->d_info) cudaFree(c->d_info);
    if (c->d_frozen_words) cudaFree(c->d_frozen_words);
    if (c->d_info_words) cudaFree(c->d_info_words);
    free(c->h_info);
    free(c);
    return NPD_OK;
}
