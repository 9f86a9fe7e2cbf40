// C-ABI glue: library identity, error text, GEMM dispatch.
#include <stdarg.h>

#include <atomic>

#include "kgb_common.cuh"

#define KGB_VERSION_MAJOR 0
#define KGB_VERSION_MINOR 1
#define KGB_VERSION_PATCH 0

namespace kgb {

char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
}

size_t gemm_ffma_workspace_bytes(int layout, int64_t m, int64_t n, int64_t k);
int gemm_ffma(int layout, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int64_t M,
              int64_t N, int64_t K, float alpha, float beta, const float* bias, int relu, void* ws, size_t ws_bytes,
              cudaStream_t stream);
bool gemm_tc_supported(int layout, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
size_t gemm_tc_workspace_bytes(int layout, int64_t m, int64_t n, int64_t k);
int gemm_tc(int layout, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc, int64_t M,
            int64_t N, int64_t K, float alpha, float beta, const float* bias, int relu, void* ws, size_t ws_bytes,
            cudaStream_t stream);

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_version(void) { return KGB_VERSION_MAJOR * 10000 + KGB_VERSION_MINOR * 100 + KGB_VERSION_PATCH; }
extern "C" int kgb_sm_arch(void) { return 100; }
extern "C" const char* kgb_last_error(void) { return err_buf(); }
extern "C" long long kgb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t kgb_gemm_workspace_bytes(int32_t layout, int64_t m, int64_t n, int64_t k) {
  size_t a = gemm_ffma_workspace_bytes(layout, m, n, k);
  size_t b = gemm_tc_workspace_bytes(layout, m, n, k);
  return a > b ? a : b;
}

extern "C" int kgb_gemm(int32_t layout, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                        int64_t m, int64_t n, int64_t k, float alpha, float beta, const float* bias, int32_t relu,
                        void* workspace, size_t workspace_bytes, kgb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  KGB_REQUIRE(layout == KGB_NT || layout == KGB_NN || layout == KGB_TN, "gemm: unknown layout %d", layout);
  KGB_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative size");
  if (m == 0 || n == 0) return KGB_OK;
  KGB_REQUIRE(a && b && c, "gemm: null pointer");
  KGB_REQUIRE(aligned16(a) && aligned16(b) && aligned16(c) && (!bias || aligned16(bias)), "gemm: 16-byte alignment");
  KGB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "gemm: strides must be multiples of 4 floats");
  KGB_REQUIRE(n % 4 == 0, "gemm: N must be a multiple of 4 (got %lld)", (long long)n);
  if (layout == KGB_TN) KGB_REQUIRE(m % 4 == 0, "gemm TN: M must be a multiple of 4");
  else KGB_REQUIRE(k % 4 == 0, "gemm NT/NN: K must be a multiple of 4");
  if (gemm_tc_supported(layout, m, n, k, lda, ldb, ldc))
    return gemm_tc(layout, a, lda, b, ldb, c, ldc, m, n, k, alpha, beta, bias, relu, workspace, workspace_bytes, stream);
  return gemm_ffma(layout, a, lda, b, ldb, c, ldc, m, n, k, alpha, beta, bias, relu, workspace, workspace_bytes, stream);
}
