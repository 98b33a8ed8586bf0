// vsb_api.cu -- error plumbing, launch accounting and host-side helpers of the C-ABI.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include "vsb_common.cuh"

static thread_local char g_err[512] = "no error";
static std::atomic<unsigned long long> g_launches{0};

void vsb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void vsb_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int vsb_check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        vsb_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return VSB_ERR_CUDA;
    }
    vsb_count_launch(1);
    return VSB_OK;
}

extern "C" const char *vsb_last_error_string(void) { return g_err; }
extern "C" int vsb_version(void) { return 100; }
extern "C" unsigned long long vsb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// T[M][m]: axis steps first (exact integers), then the diagonal steps one float32 add at a time,
// then the knight steps; symmetric in (M,m).  Same definition as oracle/c/vsb_oracle.c (written
// independently; tests compare the two tables bit for bit).
extern "C" int vsb_chamfer_table(float *T, int R)
{
    VSB_REQUIRE(T && R >= 0 && R <= 4096, "vsb_chamfer_table: bad arguments");
    const float b = 1.4f, c = 2.1969f;
    const int n = R + 1;
    for (int M = 0; M <= R; ++M)
        for (int m = 0; m <= M; ++m) {
            volatile float v;
            int nb, nc;
            if (2 * m <= M) { v = (float)(M - 2 * m); nb = 0; nc = m; }
            else { v = 0.0f; nb = 2 * m - M; nc = M - m; }
            for (int i = 0; i < nb; ++i) v = v + b;
            for (int i = 0; i < nc; ++i) v = v + c;
            T[M * n + m] = v;
            T[m * n + M] = v;
        }
    return VSB_OK;
}
