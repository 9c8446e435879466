// cab_svm.cu -- SVM classification of feature vectors (the GRSD-21 signatures) on the device.
// Replaces, per point, scaleFeature + svm_predict of cloud_algos::SVMClassification::process
// (cloud_algos/src/svm_classification.cpp:134-155; scaleFeature at
// cloud_algos/include/cloud_algos/svm_classification.h:68-86) for a libsvm C-SVC model with an RBF
// kernel (the svm/*.model files of the reference; libsvm itself is a third-party dependency,
// cloud_algos/manifest.xml:28).  svm_predict's published algorithm:
//   kvalue[s]  = exp(-gamma * sum_i (x_i - sv_s,i)^2)                      (Kernel::k_function, RBF)
//   dec(i, j)  = sum_k coef[j-1][si+k]*kvalue[si+k] + sum_k coef[i][sj+k]*kvalue[sj+k] - rho[p]
//   vote, first maximum wins, label[argmax]
// Everything is fp64 and summed in libsvm's order, so the decision values match a CPU libsvm up to
// the last bit of exp().  One block per feature vector: the kernel values are computed by all
// threads into shared memory, the class pairs are then spread over the threads and vote with
// integer shared-memory atomics.
#include <algorithm>
#include <cstring>
#include <vector>

#include "cab_internal.cuh"

namespace cab {

namespace {

constexpr int kSvmThreads = 128;

struct SvmArgs {
  const float* feat;     // n x dim (or int32 counts when from_hist)
  const int* hist;       // n x dim int32 (GRSD histograms), used when feat == nullptr
  long long n;
  int dim, nr_class, total_sv, npairs;
  double gamma, lower, upper;
  const int* labels;     // nr_class
  const int* start;      // nr_class + 1, prefix of nr_sv
  const double* rho;     // npairs
  const double* coef;    // (nr_class - 1) x total_sv
  const double* sv;      // total_sv x dim
  const double* fmin;    // dim or null
  const double* fmax;
  float* out;            // n predicted labels
  double* dec;           // optional n x npairs
};

__global__ void __launch_bounds__(kSvmThreads) svm_predict_kernel(const SvmArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* x = reinterpret_cast<double*>(smem_raw);  // [dim]
  double* kvalue = x + a.dim;                       // [total_sv]
  int* vote = reinterpret_cast<int*>(kvalue + a.total_sv);  // [nr_class]
  const long long p = blockIdx.x;
  for (int i = threadIdx.x; i < a.dim; i += blockDim.x) {
    double value = a.feat ? (double)a.feat[p * a.dim + i] : (double)(float)a.hist[p * a.dim + i];
    if (a.fmin) {  // scaleFeature
      const double lo = a.fmin[i], hi = a.fmax[i];
      if (lo == hi) value = 0;
      else if (value <= lo) value = a.lower;
      else if (value >= hi) value = a.upper;
      else value = __dadd_rn(a.lower, __ddiv_rn(__dmul_rn(__dsub_rn(a.upper, a.lower), __dsub_rn(value, lo)), __dsub_rn(hi, lo)));
    }
    x[i] = value;
  }
  for (int i = threadIdx.x; i < a.nr_class; i += blockDim.x) vote[i] = 0;
  __syncthreads();
  for (int s = threadIdx.x; s < a.total_sv; s += blockDim.x) {
    const double* y = a.sv + (size_t)s * a.dim;
    double sum = 0;
    for (int i = 0; i < a.dim; ++i) {
      const double d = __dsub_rn(x[i], y[i]);
      sum = __dadd_rn(sum, __dmul_rn(d, d));
    }
    kvalue[s] = exp(__dmul_rn(-a.gamma, sum));
  }
  __syncthreads();
  for (int q = threadIdx.x; q < a.npairs; q += blockDim.x) {
    // pair index q -> (i, j), i < j, in libsvm's loop order
    int i = 0, rem = q;
    while (rem >= a.nr_class - 1 - i) {
      rem -= a.nr_class - 1 - i;
      ++i;
    }
    const int j = i + 1 + rem;
    const int si = a.start[i], sj = a.start[j], ci = a.start[i + 1] - si, cj = a.start[j + 1] - sj;
    const double* coef1 = a.coef + (size_t)(j - 1) * a.total_sv;
    const double* coef2 = a.coef + (size_t)i * a.total_sv;
    double sum = 0;
    for (int k = 0; k < ci; ++k) sum = __dadd_rn(sum, __dmul_rn(coef1[si + k], kvalue[si + k]));
    for (int k = 0; k < cj; ++k) sum = __dadd_rn(sum, __dmul_rn(coef2[sj + k], kvalue[sj + k]));
    sum = __dsub_rn(sum, a.rho[q]);
    if (a.dec) a.dec[p * a.npairs + q] = sum;
    atomicAdd(&vote[sum > 0 ? i : j], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int i = 1; i < a.nr_class; ++i)
      if (vote[i] > vote[best]) best = i;
    a.out[p] = (float)a.labels[best];
  }
}

}  // namespace

struct SvmState {
  int dim = 0, nr_class = 0, total_sv = 0;
  double gamma = 0, lower = 0, upper = 0;
  bool have_model = false, have_scale = false;
  DevBuf labels, start, rho, coef, sv, fmin, fmax, feat, out, dec;
};

static int upload(cab_ctx* ctx, DevBuf& b, const void* src, size_t bytes) {
  if (int rc = reserve(ctx, b, std::max<size_t>(bytes, 16))) return rc;
  if (bytes) CAB_CUDA(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // src may be pageable and short-lived
  return CAB_OK;
}

static SvmState* state(cab_ctx* ctx) {
  if (!ctx->svm) ctx->svm = new SvmState();
  return ctx->svm;
}

void svm_free(cab_ctx* ctx) {
  if (!ctx->svm) return;
  SvmState* s = ctx->svm;
  DevBuf* bufs[] = {&s->labels, &s->start, &s->rho, &s->coef, &s->sv, &s->fmin, &s->fmax, &s->feat, &s->out, &s->dec};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  delete s;
  ctx->svm = nullptr;
}

static int run_predict(cab_ctx* ctx, const float* d_feat, const int* d_hist, int64_t n, float* point_class, double* dec_values) {
  SvmState* s = ctx->svm;
  const int npairs = s->nr_class * (s->nr_class - 1) / 2;
  if (int rc = reserve(ctx, s->out, (size_t)std::max<int64_t>(n, 1) * sizeof(float))) return rc;
  if (dec_values)
    if (int rc = reserve(ctx, s->dec, (size_t)std::max<int64_t>(n, 1) * npairs * sizeof(double))) return rc;
  if (n > 0) {
    SvmArgs a{};
    a.feat = d_feat;
    a.hist = d_hist;
    a.n = n;
    a.dim = s->dim;
    a.nr_class = s->nr_class;
    a.total_sv = s->total_sv;
    a.npairs = npairs;
    a.gamma = s->gamma;
    a.lower = s->lower;
    a.upper = s->upper;
    a.labels = (const int*)s->labels.p;
    a.start = (const int*)s->start.p;
    a.rho = (const double*)s->rho.p;
    a.coef = (const double*)s->coef.p;
    a.sv = (const double*)s->sv.p;
    a.fmin = s->have_scale ? (const double*)s->fmin.p : nullptr;
    a.fmax = s->have_scale ? (const double*)s->fmax.p : nullptr;
    a.out = (float*)s->out.p;
    a.dec = dec_values ? (double*)s->dec.p : nullptr;
    const size_t smem = (size_t)(s->dim + s->total_sv) * sizeof(double) + (size_t)s->nr_class * sizeof(int);
    if (smem > 200 * 1024) return fail(ctx, CAB_ERR_ARG, "cab_svm_predict: model too large for shared memory (%zu bytes)", smem);
    CAB_CUDA(ctx, cudaFuncSetAttribute(svm_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    svm_predict_kernel<<<(unsigned)n, kSvmThreads, smem, ctx->stream>>>(a);
    CAB_LAUNCH_CHECK(ctx);
    if (point_class)
      CAB_CUDA(ctx, cudaMemcpyAsync(point_class, s->out.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (dec_values)
      CAB_CUDA(ctx, cudaMemcpyAsync(dec_values, s->dec.p, (size_t)n * npairs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_svm_set_model(cab_ctx* ctx, int32_t dim, int32_t nr_class, int32_t total_sv, double gamma, const int32_t* labels,
                      const int32_t* nr_sv, const double* rho, const double* sv_coef, const double* sv) {
  if (!ctx) return CAB_ERR_ARG;
  if (dim < 1 || nr_class < 2 || total_sv < 1 || !labels || !nr_sv || !rho || !sv_coef || !sv)
    return fail(ctx, CAB_ERR_ARG, "cab_svm_set_model: bad model (dim %d, nr_class %d, total_sv %d)", dim, nr_class, total_sv);
  std::vector<int> start(nr_class + 1, 0);
  for (int i = 0; i < nr_class; ++i) {
    if (nr_sv[i] < 0) return fail(ctx, CAB_ERR_ARG, "cab_svm_set_model: negative nr_sv");
    start[i + 1] = start[i] + nr_sv[i];
  }
  if (start[nr_class] != total_sv) return fail(ctx, CAB_ERR_ARG, "cab_svm_set_model: nr_sv does not sum to total_sv");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  SvmState* s = state(ctx);
  s->have_model = false;
  if (s->dim != dim) s->have_scale = false;
  const int npairs = nr_class * (nr_class - 1) / 2;
  if (int rc = upload(ctx, s->labels, labels, (size_t)nr_class * 4)) return rc;
  if (int rc = upload(ctx, s->start, start.data(), (size_t)(nr_class + 1) * 4)) return rc;
  if (int rc = upload(ctx, s->rho, rho, (size_t)npairs * 8)) return rc;
  if (int rc = upload(ctx, s->coef, sv_coef, (size_t)(nr_class - 1) * total_sv * 8)) return rc;
  if (int rc = upload(ctx, s->sv, sv, (size_t)total_sv * dim * 8)) return rc;
  s->dim = dim;
  s->nr_class = nr_class;
  s->total_sv = total_sv;
  s->gamma = gamma;
  s->have_model = true;
  return CAB_OK;
}

int cab_svm_set_scaling(cab_ctx* ctx, int32_t dim, double lower, double upper, const double* fmin, const double* fmax) {
  if (!ctx) return CAB_ERR_ARG;
  SvmState* s = state(ctx);
  if (!fmin && !fmax) {
    s->have_scale = false;
    return CAB_OK;
  }
  if (!fmin || !fmax || dim < 1) return fail(ctx, CAB_ERR_ARG, "cab_svm_set_scaling: give both fmin and fmax");
  if (s->have_model && dim != s->dim) return fail(ctx, CAB_ERR_ARG, "cab_svm_set_scaling: %d ranges for a %d-dimensional model", dim, s->dim);
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = upload(ctx, s->fmin, fmin, (size_t)dim * 8)) return rc;
  if (int rc = upload(ctx, s->fmax, fmax, (size_t)dim * 8)) return rc;
  s->lower = lower;
  s->upper = upper;
  s->have_scale = true;
  if (!s->have_model) s->dim = dim;
  return CAB_OK;
}

int cab_svm_predict(cab_ctx* ctx, const float* features, int64_t n, int32_t dim, float* point_class, double* dec_values) {
  if (!ctx) return CAB_ERR_ARG;
  SvmState* s = ctx->svm;
  if (!s || !s->have_model) return fail(ctx, CAB_ERR_STATE, "cab_svm_predict: no model (cab_svm_set_model)");
  if (dim != s->dim) return fail(ctx, CAB_ERR_ARG, "cab_svm_predict: %d features, the model has %d", dim, s->dim);
  if (n < 0 || (n > 0 && !features)) return fail(ctx, CAB_ERR_ARG, "cab_svm_predict: bad features");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = reserve(ctx, s->feat, (size_t)std::max<int64_t>(n, 1) * dim * sizeof(float))) return rc;
  if (n > 0) CAB_CUDA(ctx, cudaMemcpyAsync(s->feat.p, features, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return run_predict(ctx, (const float*)s->feat.p, nullptr, n, point_class, dec_values);
}

int cab_svm_predict_grsd(cab_ctx* ctx, float* point_class) {
  if (!ctx) return CAB_ERR_ARG;
  SvmState* s = ctx->svm;
  if (!s || !s->have_model) return fail(ctx, CAB_ERR_STATE, "cab_svm_predict_grsd: no model (cab_svm_set_model)");
  if (s->dim != 21) return fail(ctx, CAB_ERR_ARG, "cab_svm_predict_grsd: the model has %d features, GRSD has 21", s->dim);
  const int nd = (int)ctx->g_min_div.size() / 6;
  if (nd == 0) return fail(ctx, CAB_ERR_STATE, "cab_svm_predict_grsd: run cab_grsd_batch first");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_predict(ctx, nullptr, (const int*)ctx->g_hist.p, nd, point_class, nullptr);
}

}  // extern "C"
