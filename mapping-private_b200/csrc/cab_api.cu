// cab_api.cu -- C ABI of libcloudalgos_b200.so (see include/cloud_algos_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <mutex>
#include <numeric>

#include "cab_internal.cuh"

namespace cab {

static thread_local std::string g_create_err;

int fail(cab_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}

int reserve(cab_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return CAB_OK;
  if (b.p) {
    // The outgrown buffer is not freed here: cudaFree synchronises the whole device, and in a group whose ranks share a
    // GPU (several contexts of one process, the tests) a peer's kernel may be spinning on a flag this rank has yet to
    // post -- the free would wait for the peer and the peer for this rank.  Work still in flight may also go on using
    // the old buffer.  The graveyard is emptied when the context is destroyed, or here once it holds more than a GiB.
    ctx->graveyard.push_back(b);
    ctx->graveyard_bytes += b.cap;
    b.p = nullptr;
    b.cap = 0;
    if (ctx->graveyard_bytes > ((size_t)1 << 30)) {
      cudaStreamSynchronize(ctx->stream);
      cudaStreamSynchronize(ctx->copy_stream);
      for (DevBuf& g : ctx->graveyard) cudaFree(g.p);
      ctx->graveyard.clear();
      ctx->graveyard_bytes = 0;
    }
  }
  size_t want = bytes + bytes / 8 + 256;  // grow-only arena with a little headroom
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    b.p = nullptr;
    return fail(ctx, CAB_ERR_OOM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  return CAB_OK;
}

int reserve_pinned(cab_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_pin_cap) return CAB_OK;
  if (ctx->h_pin) {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->h_pin_cap = 0;
  }
  size_t want = std::max<size_t>(bytes * 2, 1 << 16);
  cudaError_t e = cudaMallocHost(&ctx->h_pin, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, CAB_ERR_OOM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  ctx->h_pin_cap = want;
  return CAB_OK;
}

void read_stats(cab_ctx* ctx, const void* staged) {
  const unsigned long long* h = (const unsigned long long*)(staged ? staged : ctx->h_pin);
  unsigned long long ks = 0, cs = 0;
  for (int i = 0; i < kStatSlots; ++i) {
    ks += h[2 * i];
    cs += h[2 * i + 1];
  }
  ctx->tm.neighbour_sum = (int64_t)ks;
  ctx->tm.candidate_sum = (int64_t)cs;
}

namespace {

__global__ void __launch_bounds__(256) gather_normals_kernel(const float* __restrict__ nx, const float* __restrict__ ny,
                                                             const float* __restrict__ nz, const int* __restrict__ perm,
                                                             int n, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int j = perm[i];
  if (j < 0) return;  // position not held by this shard (sharded sort)
  out[i] = make_float4(nx[j], ny[j], nz[j], 0.f);
}

__global__ void __launch_bounds__(256) unpermute_kernel(const int* __restrict__ perm, int begin, int end,
                                                        const float4* __restrict__ nrm, const float2* __restrict__ rsd,
                                                        float4* __restrict__ out4, float* __restrict__ out_a,
                                                        float* __restrict__ out_b) {
  int i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= end) return;
  int j = perm[i];
  if (j < 0) return;  // position not held by this shard (sharded sort)
  if (out4) out4[j] = nrm[i];
  if (out_a) {
    float2 v = rsd[i];
    out_a[j] = v.x;
    out_b[j] = v.y;
  }
}

__global__ void __launch_bounds__(256) unpermute_scalar_kernel(const int* __restrict__ perm, int count,
                                                               const float* __restrict__ in, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int j = perm[i];
  if (j >= 0) out[j] = in[i];
}

int set_domains(cab_ctx* ctx, int64_t n, const int32_t* offsets, int nclusters) {
  ctx->dom_offsets.clear();
  if (offsets) {
    if (nclusters < 1) return fail(ctx, CAB_ERR_ARG, "nclusters must be >= 1");
    if (offsets[0] != 0 || offsets[nclusters] != n) return fail(ctx, CAB_ERR_ARG, "offsets must span [0, n]");
    for (int c = 0; c < nclusters; ++c)
      if (offsets[c + 1] < offsets[c]) return fail(ctx, CAB_ERR_ARG, "offsets must be non-decreasing");
    ctx->dom_offsets.assign(offsets, offsets + nclusters + 1);
  } else {
    ctx->dom_offsets = {0, (int32_t)n};
  }
  ctx->n_domains = (int)ctx->dom_offsets.size() - 1;
  size_t bytes = ctx->dom_offsets.size() * sizeof(int32_t);
  if (int rc = reserve(ctx, ctx->b_domoff, bytes)) return rc;
  if (int rc = reserve_pinned(ctx, bytes)) return rc;
  std::memcpy(ctx->h_pin, ctx->dom_offsets.data(), bytes);
  CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_domoff.p, ctx->h_pin, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CAB_OK;
}

int upload_impl(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride, const int32_t* offsets, int nclusters,
                bool device_ptr) {
  if (!ctx) return CAB_ERR_ARG;
  if (n < 0 || n > 0x7ffffff0ll) return fail(ctx, CAB_ERR_ARG, "n out of range");
  if (stride < 3) return fail(ctx, CAB_ERR_ARG, "stride must be >= 3 floats");
  if (n > 0 && !xyz) return fail(ctx, CAB_ERR_ARG, "xyz is NULL");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->have_cloud = ctx->have_grid = ctx->have_normals = ctx->have_rsd = ctx->kcount_valid = ctx->trunc_hist_valid = false;
  ctx->g_min_div.clear();  // voxel state of the last cab_grsd_batch belongs to the previous cloud
  if (int rc = set_domains(ctx, n, offsets, nclusters)) return rc;
  if (device_ptr) {
    ctx->xyz_in = xyz;
  } else {
    size_t bytes = (size_t)n * stride * sizeof(float);
    if (int rc = reserve(ctx, ctx->b_xyz, std::max<size_t>(bytes, 16))) return rc;
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
    if (bytes) CAB_CUDA(ctx, cudaMemcpyAsync(ctx->b_xyz.p, xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
    CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[6], ctx->ev[7]));
    ctx->xyz_in = (const float*)ctx->b_xyz.p;
  }
  ctx->cloud_external = device_ptr;
  ctx->n = n;
  ctx->stride = stride;
  ctx->have_cloud = true;
  return CAB_OK;
}

}  // namespace

int permute_normals_in(cab_ctx* ctx, const float* nx, const float* ny, const float* nz) {
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream;
  if (int rc = reserve(ctx, ctx->b_nrm_in, (size_t)std::max(n, 1) * 3 * sizeof(float))) return rc;
  if (int rc = reserve(ctx, ctx->b_nrm, (size_t)std::max(n, 1) * sizeof(float4))) return rc;
  float* d = (float*)ctx->b_nrm_in.p;
  if (n > 0) {
    CAB_CUDA(ctx, cudaMemcpyAsync(d, nx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CAB_CUDA(ctx, cudaMemcpyAsync(d + n, ny, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CAB_CUDA(ctx, cudaMemcpyAsync(d + 2 * (size_t)n, nz, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    const int ns = (int)ctx->n_sorted;
    if (ns > 0) {
      gather_normals_kernel<<<(ns + 255) / 256, 256, 0, st>>>(d, d + n, d + 2 * (size_t)n, (const int*)ctx->b_perm.p, ns,
                                                              (float4*)ctx->b_nrm.p);
      CAB_LAUNCH_CHECK(ctx);
    }
  }
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->have_normals = true;
  return CAB_OK;
}

int download_results(cab_ctx* ctx, float* n4, float* rmin, float* rmax) {
  const int n = (int)ctx->n;
  const int ns = (int)ctx->n_sorted;  // entries of the sorted arrays (a slab holds fewer than n)
  cudaStream_t st = ctx->stream;
  if ((rmin == nullptr) != (rmax == nullptr)) return fail(ctx, CAB_ERR_ARG, "r_min and r_max must be given together");
  if (n4 && !ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "no normals to download");
  if (rmin && !ctx->have_rsd) return fail(ctx, CAB_ERR_STATE, "no RSD results to download");
  if (n == 0 || (!n4 && !rmin)) return CAB_OK;
  if (n4)
    if (int rc = reserve(ctx, ctx->b_out4, (size_t)n * sizeof(float4))) return rc;
  if (rmin) {
    if (int rc = reserve(ctx, ctx->b_out1a, (size_t)n * sizeof(float))) return rc;
    if (int rc = reserve(ctx, ctx->b_out1b, (size_t)n * sizeof(float))) return rc;
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
  if (ctx->slab) {  // rows of other shards: untouched device memory must not reach the caller
    if (n4) CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_out4.p, 0, (size_t)n * sizeof(float4), st));
    if (rmin) {
      CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_out1a.p, 0, (size_t)n * sizeof(float), st));
      CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_out1b.p, 0, (size_t)n * sizeof(float), st));
    }
  }
  if (ns > 0) {
    unpermute_kernel<<<(ns + 255) / 256, 256, 0, st>>>((const int*)ctx->b_perm.p, 0, ns, (const float4*)ctx->b_nrm.p,
                                                     (const float2*)ctx->b_rsd.p, n4 ? (float4*)ctx->b_out4.p : nullptr,
                                                     rmin ? (float*)ctx->b_out1a.p : nullptr,
                                                     rmin ? (float*)ctx->b_out1b.p : nullptr);
    CAB_LAUNCH_CHECK(ctx);
  }
  if (n4) CAB_CUDA(ctx, cudaMemcpyAsync(n4, ctx->b_out4.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, st));
  if (rmin) {
    CAB_CUDA(ctx, cudaMemcpyAsync(rmin, ctx->b_out1a.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CAB_CUDA(ctx, cudaMemcpyAsync(rmax, ctx->b_out1b.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[6], ctx->ev[7]));
  return CAB_OK;
}

}  // namespace cab

using namespace cab;

extern "C" {

int cab_version(void) { return 100; }

int cab_create(const cab_config* cfg, cab_ctx** out) {
  if (!out) return CAB_ERR_ARG;
  *out = nullptr;
  cab_config c{};
  if (cfg) c = *cfg;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, CAB_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (c.device < 0 || c.device >= ndev) return fail(nullptr, CAB_ERR_ARG, "device %d out of range [0,%d)", c.device, ndev);
  cudaDeviceProp prop{};
  if ((e = cudaGetDeviceProperties(&prop, c.device)) != cudaSuccess)
    return fail(nullptr, CAB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, CAB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", c.device,
                prop.major, prop.minor);
  if ((e = cudaSetDevice(c.device)) != cudaSuccess)
    return fail(nullptr, CAB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cab_ctx* ctx = new cab_ctx();
  ctx->cfg = c;
  ctx->device = c.device;
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* hp = std::getenv("CAB_HALO_PERMILLE")) {  // shard balance knob (tuning runs only)
    const int v = std::atoi(hp);
    if (v >= 0 && v <= 2000) ctx->halo_permille = v;
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, CAB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return fail(nullptr, CAB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (cudaMallocHost((void**)&ctx->h_step, kStepBytes) != cudaSuccess) {
    cudaGetLastError();
    g_create_err = "cudaMallocHost failed";
    cab_destroy(ctx);
    return CAB_ERR_OOM;
  }
  if (reserve_pinned(ctx, 1 << 16) != CAB_OK) {
    g_create_err = ctx->err;
    cab_destroy(ctx);
    return CAB_ERR_OOM;
  }
  *out = ctx;
  return CAB_OK;
}

void cab_destroy(cab_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  comm_free(ctx);  // first: while the context belongs to a group its working normals live in a buffer the group owns
  DevBuf* bufs[] = {&ctx->b_xyz, &ctx->b_domoff, &ctx->b_domid, &ctx->b_bounds, &ctx->b_domains, &ctx->b_keys[0],
                    &ctx->b_keys[1], &ctx->b_keys[2], &ctx->b_vals[0], &ctx->b_vals[1], &ctx->b_vals[2], &ctx->b_thr_flag, &ctx->b_knn_avg, &ctx->b_knn_done, &ctx->b_pfh[0], &ctx->b_pfh[1], &ctx->b_pfh[2], &ctx->b_cluster, &ctx->b_cubtmp, &ctx->b_pos, &ctx->b_perm,
                    &ctx->b_cellcnt, &ctx->b_cellstart, &ctx->b_rowpk, &ctx->b_packets, &ctx->b_nrm, &ctx->b_nrm_in,
                    &ctx->b_rsd, &ctx->b_rdif, &ctx->b_kcount, &ctx->b_stats, &ctx->b_out4, &ctx->b_out1a, &ctx->b_out1b,
                    &ctx->b_thr_d2, &ctx->b_thr_idx, &ctx->b_misc, &ctx->b_pcost, &ctx->b_slab, &ctx->b_sel, &ctx->b_stats2, &ctx->b_sorttmp, &ctx->b_occ, &ctx->b_in_nrm, &ctx->b_in_rsd, &ctx->g_vkeys[0], &ctx->g_vkeys[1], &ctx->g_vvals[0],
                    &ctx->g_vvals[1], &ctx->g_cent, &ctx->g_vcount, &ctx->g_vrad, &ctx->g_vlabel, &ctx->g_voff,
                    &ctx->g_layout, &ctx->g_layoff, &ctx->g_vgrid, &ctx->g_hist, &ctx->g_vfirst, &ctx->g_cnrm, &ctx->g_invperm, &ctx->g_sig, &ctx->g_sigdom, &ctx->g_color, &ctx->g_vown};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (DevBuf& g : ctx->graveyard) cudaFree(g.p);
  svm_free(ctx);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->h_step) cudaFreeHost(ctx->h_step);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* cab_last_error(const cab_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int cab_upload_cloud(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride) {
  return upload_impl(ctx, xyz, n, stride, nullptr, 0, false);
}

int cab_upload_clusters(cab_ctx* ctx, const float* xyz, int64_t n, int32_t stride, const int32_t* offsets,
                        int32_t nclusters) {
  if (!ctx) return CAB_ERR_ARG;
  if (!offsets) return fail(ctx, CAB_ERR_ARG, "offsets is NULL");
  return upload_impl(ctx, xyz, n, stride, offsets, nclusters, false);
}

int cab_set_cloud_device(cab_ctx* ctx, const float* d_xyz, int64_t n, int32_t stride) {
  return upload_impl(ctx, d_xyz, n, stride, nullptr, 0, true);
}

int cab_build_grid(cab_ctx* ctx, float cell) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return build_grid(ctx, cell);
}

int cab_set_shard(cab_ctx* ctx, int32_t rank, int32_t world) {
  if (!ctx) return CAB_ERR_ARG;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, CAB_ERR_ARG, "bad shard %d/%d", rank, world);
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  return CAB_OK;
}

int cab_shard_range(cab_ctx* ctx, int64_t* begin, int64_t* end) {
  if (!ctx || !begin || !end) return CAB_ERR_ARG;
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_shard_range: no grid");
  if (!ctx->slab) {
    *begin = 0;
    *end = ctx->n_valid;
    return CAB_OK;
  }
  if (!ctx->slab_info_valid) {
    CAB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int rc = finish_slab(ctx)) return rc;
  }
  *begin = ctx->slab_info.q0;
  *end = ctx->slab_info.q1;
  return CAB_OK;
}

int cab_normals(cab_ctx* ctx, float r, int32_t max_nn, const float vp[3], float* nxyz_curv) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = run_normals(ctx, r, max_nn, vp)) return rc;
  if (nxyz_curv) return download_results(ctx, nxyz_curv, nullptr, nullptr);
  return CAB_OK;
}

int cab_set_normals(cab_ctx* ctx, const float* nx, const float* ny, const float* nz) {
  if (!ctx) return CAB_ERR_ARG;
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_set_normals: build the grid first");
  if (ctx->n > 0 && (!nx || !ny || !nz)) return fail(ctx, CAB_ERR_ARG, "cab_set_normals: NULL channel");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return permute_normals_in(ctx, nx, ny, nz);
}

int cab_rsd(cab_ctx* ctx, double r, int32_t max_nn, int32_t ndiv, double plane_radius, int32_t flags, float* r_min,
            float* r_max) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((r_min == nullptr) != (r_max == nullptr)) return fail(ctx, CAB_ERR_ARG, "r_min and r_max must be given together");
  if (int rc = run_rsd(ctx, r, max_nn, ndiv, plane_radius, flags)) return rc;
  if (r_min) return download_results(ctx, nullptr, r_min, r_max);
  return CAB_OK;
}

int64_t cab_neighbors_debug(cab_ctx* ctx, float r, int32_t max_nn, int64_t q0, int64_t q1, int64_t* offsets,
                            int32_t* idx, float* d2, int64_t cap) {
  if (!ctx) return CAB_ERR_ARG;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cudaSetDevice failed");
  return run_neighbors_debug(ctx, r, max_nn, q0, q1, offsets, idx, d2, cap);
}

void* cab_device_ptr(cab_ctx* ctx, int32_t which) {
  if (!ctx) return nullptr;
  switch (which) {
    case CAB_BUF_POS_SORTED: return ctx->b_pos.p;
    case CAB_BUF_NRM_SORTED: return ctx->b_nrm.p;
    case CAB_BUF_RSD_SORTED: return ctx->b_rsd.p;
    case CAB_BUF_PERM: return ctx->b_perm.p;
    case CAB_BUF_NRM_INPUT_RANGE: return ctx->have_input_order ? ctx->b_in_nrm.p : nullptr;  // CAB_STEP_INPUT_ORDER
    case CAB_BUF_RSD_INPUT_RANGE: return ctx->have_input_order ? ctx->b_in_rsd.p : nullptr;
    default: return nullptr;
  }
}

void* cab_stream(cab_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int cab_download(cab_ctx* ctx, float* nxyz_curv, float* r_min, float* r_max) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  return download_results(ctx, nxyz_curv, r_min, r_max);
}

int cab_download_rdif(cab_ctx* ctx, float* r_dif) {
  if (!ctx || !r_dif) return CAB_ERR_ARG;
  if (!ctx->have_rsd) return fail(ctx, CAB_ERR_STATE, "cab_download_rdif: no RSD results");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = (int)ctx->n, ns = (int)ctx->n_sorted;
  if (n == 0) return CAB_OK;
  cudaStream_t st = ctx->stream;
  if (int rc = reserve(ctx, ctx->b_out1a, (size_t)n * sizeof(float))) return rc;
  CAB_CUDA(ctx, cudaMemsetAsync(ctx->b_out1a.p, 0, (size_t)n * sizeof(float), st));
  if (ns > 0) {
    unpermute_scalar_kernel<<<(ns + 255) / 256, 256, 0, st>>>((const int*)ctx->b_perm.p, ns, (const float*)ctx->b_rdif.p,
                                                            (float*)ctx->b_out1a.p);
    CAB_LAUNCH_CHECK(ctx);
  }
  CAB_CUDA(ctx, cudaMemcpyAsync(r_dif, ctx->b_out1a.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  return CAB_OK;
}

int cab_download_sorted(cab_ctx* ctx, int64_t begin, int64_t end, float* nxyz_curv, float* rmin_rmax, int32_t* input_index) {
  if (!ctx) return CAB_ERR_ARG;
  if (!ctx->have_grid) return fail(ctx, CAB_ERR_STATE, "cab_download_sorted: no grid");
  if (begin < 0 || end < begin || end > ctx->n_sorted) return fail(ctx, CAB_ERR_ARG, "cab_download_sorted: bad range");
  if (nxyz_curv && !ctx->have_normals) return fail(ctx, CAB_ERR_STATE, "no normals to download");
  if (rmin_rmax && !ctx->have_rsd) return fail(ctx, CAB_ERR_STATE, "no RSD results to download");
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t m = (size_t)(end - begin);
  cudaStream_t st = ctx->stream;
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
  if (m && nxyz_curv)
    CAB_CUDA(ctx, cudaMemcpyAsync(nxyz_curv, (const float4*)ctx->b_nrm.p + begin, m * sizeof(float4), cudaMemcpyDeviceToHost, st));
  if (m && rmin_rmax)
    CAB_CUDA(ctx, cudaMemcpyAsync(rmin_rmax, (const float2*)ctx->b_rsd.p + begin, m * sizeof(float2), cudaMemcpyDeviceToHost, st));
  if (m && input_index)
    CAB_CUDA(ctx, cudaMemcpyAsync(input_index, (const int*)ctx->b_perm.p + begin, m * sizeof(int), cudaMemcpyDeviceToHost, st));
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
  CAB_CUDA(ctx, cudaStreamSynchronize(st));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[6], ctx->ev[7]));
  return CAB_OK;
}

int cab_normals_rsd(cab_ctx* ctx, double r, int32_t max_nn_normals, const float vp[3], int32_t max_nn_rsd, int32_t ndiv,
                    double plane_radius, int32_t flags, int32_t layout, float* nxyz_curv, float* out_a, float* out_b,
                    int32_t* input_index) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (layout != CAB_OUT_INPUT_ORDER && layout != CAB_OUT_SHARD_SORTED)
    return fail(ctx, CAB_ERR_ARG, "cab_normals_rsd: unknown output layout %d", layout);
  if (layout == CAB_OUT_INPUT_ORDER && (out_a == nullptr) != (out_b == nullptr))
    return fail(ctx, CAB_ERR_ARG, "cab_normals_rsd: r_min and r_max must be given together");
  const int n = (int)ctx->n;
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  if (layout == CAB_OUT_INPUT_ORDER && ctx->slab)
    return fail(ctx, CAB_ERR_STATE, "cab_normals_rsd: a shard returns its own slice (CAB_OUT_SHARD_SORTED); the concatenated "
                                     "results of a group come from cab_step_normals_rsd + cab_comm_download_range");
  // Input-order layout: both pass kernels store every query's result at its input index as well (one scattered store
  // per query next to the sorted-order one), so the device-to-host copies leave straight behind the kernels -- no
  // permutation pass in between.
  struct Fused {
    cab_ctx* c;
    explicit Fused(cab_ctx* x) : c(x) {}
    ~Fused() { c->fuse_nrm_in = nullptr, c->fuse_rmin_in = c->fuse_rmax_in = nullptr; }
  } fused(ctx);
  if (layout == CAB_OUT_INPUT_ORDER && n > 0) {
    if (nxyz_curv) {
      if (int rc = reserve(ctx, ctx->b_out4, (size_t)n * sizeof(float4))) return rc;
      ctx->fuse_nrm_in = (float4*)ctx->b_out4.p;
    }
    if (out_a) {
      if (int rc = reserve(ctx, ctx->b_out1a, (size_t)n * sizeof(float))) return rc;
      if (int rc = reserve(ctx, ctx->b_out1b, (size_t)n * sizeof(float))) return rc;
    }
  }
  // pass 1 on the compute stream
  // (an RSD pass truncated at max_nn_rsd follows at the same radius: the normals traversal takes its d2 histogram along)
  if (int rc = run_normals(ctx, (float)r, max_nn_normals, vp, nullptr, max_nn_normals == 0 ? max_nn_rsd : 0)) return rc;
  ctx->fuse_nrm_in = nullptr;
  int64_t b = 0, e = 0;
  if (layout == CAB_OUT_SHARD_SORTED)
    if (int rc = cab_shard_range(ctx, &b, &e)) return rc;
  const size_t m = (size_t)(e - b);
  // From here on the copy stream may be writing into the caller's buffers: every exit goes through the epilogue below,
  // which drains both streams; the first error (a failing call, or a CUDA error of any enqueue) is the result.
  int rc = CAB_OK;
  cudaError_t ce = cudaSuccess;
  auto cuda = [&](cudaError_t x) {
    if (ce == cudaSuccess && x != cudaSuccess) ce = x;
    return x == cudaSuccess;
  };
  // the normals leave on the copy stream while pass 2 runs
  cuda(cudaEventRecord(ctx->ev[6], st));
  if (nxyz_curv && n > 0) {
    if (layout == CAB_OUT_INPUT_ORDER) {  // b_out4 was filled by the normals kernel itself
      if (cuda(cudaEventRecord(ctx->ev_ready, st)) && cuda(cudaStreamWaitEvent(cs, ctx->ev_ready, 0)))
        cuda(cudaMemcpyAsync(nxyz_curv, ctx->b_out4.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, cs));
    } else if (m) {
      if (cuda(cudaEventRecord(ctx->ev_ready, st)) && cuda(cudaStreamWaitEvent(cs, ctx->ev_ready, 0)))
        cuda(cudaMemcpyAsync(nxyz_curv, (const float4*)ctx->b_nrm.p + b, m * sizeof(float4), cudaMemcpyDeviceToHost, cs));
    }
  }
  if (ce == cudaSuccess && layout == CAB_OUT_SHARD_SORTED && input_index && m)
    cuda(cudaMemcpyAsync(input_index, (const int*)ctx->b_perm.p + b, m * sizeof(int), cudaMemcpyDeviceToHost, cs));
  // pass 2 (run_rsd synchronises the compute stream only)
  if (layout == CAB_OUT_INPUT_ORDER && out_a && n > 0) {
    ctx->fuse_rmin_in = (float*)ctx->b_out1a.p;
    ctx->fuse_rmax_in = (float*)ctx->b_out1b.p;
  }
  if (ce == cudaSuccess) rc = run_rsd(ctx, r, max_nn_rsd, ndiv, plane_radius, flags);
  if (rc == CAB_OK && ce == cudaSuccess && out_a && n > 0) {
    if (layout == CAB_OUT_INPUT_ORDER) {  // b_out1a / b_out1b were filled by the RSD kernel itself
      cuda(cudaMemcpyAsync(out_a, ctx->b_out1a.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
      cuda(cudaMemcpyAsync(out_b, ctx->b_out1b.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else if (m) {
      cuda(cudaMemcpyAsync(out_a, (const float2*)ctx->b_rsd.p + b, m * sizeof(float2), cudaMemcpyDeviceToHost, st));
    }
  }
  cuda(cudaEventRecord(ctx->ev[7], st));
  // epilogue: both streams drain before the host buffers are valid or reusable, on every path
  cuda(cudaStreamSynchronize(cs));
  cuda(cudaStreamSynchronize(st));
  cuda(cudaGetLastError());
  if (rc != CAB_OK) return rc;
  if (ce != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cab_normals_rsd: %s", cudaGetErrorString(ce));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[6], ctx->ev[7]));
  ctx->tm.d2h_ms -= ctx->tm.rsd_ms;  // copy time left exposed around pass 2
  return CAB_OK;
}

int cab_step_normals_rsd(cab_ctx* ctx, float cell, double r, int32_t max_nn_normals, const float vp[3], int32_t max_nn_rsd,
                         int32_t ndiv, double plane_radius, int32_t flags) {
  if (!ctx) return CAB_ERR_ARG;
  CAB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const bool exchange = comm_active(ctx);
  struct Deferred {  // the stages leave the stream running; whatever happens, the flag goes back
    cab_ctx* c;
    explicit Deferred(cab_ctx* x) : c(x) { c->defer_sync = true; }
    ~Deferred() { c->defer_sync = false; }
  } deferred(ctx);
  CAB_CUDA(ctx, cudaEventRecord(ctx->ev[8], st));
  int rc = CAB_OK;
  ctx->have_input_order = false;
  struct InputOrder {  // the RSD pass of this call scatters its results into input-order arrays (outside a group)
    cab_ctx* c;
    InputOrder(cab_ctx* x, bool on) : c(x) { c->step_input_order = on; }
    ~InputOrder() { c->step_input_order = false; }
  } input_order(ctx, (flags & CAB_STEP_INPUT_ORDER) != 0 && !exchange && ctx->shard_world == 1);
  flags &= ~CAB_STEP_INPUT_ORDER;
  if (exchange) rc = comm_prepare(ctx, ctx->n);  // before the build: the working normals move into the exported buffer
  ctx->want_halo_exchange = exchange && rc == CAB_OK;
  if (ctx->want_halo_exchange) {
    comm_range_defaults(ctx, plane_radius);
    comm_shares(ctx);
  }
  if (rc == CAB_OK) rc = build_grid(ctx, cell);
  ctx->want_halo_exchange = false;
  ctx->range_defaults = RangeDefaults{};
  ctx->shard_cum.clear();
  const bool push = rc == CAB_OK && exchange && ctx->slab;
  const bool halo = push && ctx->slab_info.exchange != 0;  // known since the build's second host round trip
  if (rc == CAB_OK && exchange && !ctx->slab && ctx->n_valid > 0)
    rc = fail(ctx, CAB_ERR_STATE, "cab_step_normals_rsd: the group's shard was reset (cab_set_shard) behind its back");
  if (rc == CAB_OK && push) {
    // every buffer the two passes need, before the first kernel that waits for a peer is enqueued: an allocation behind
    // such a kernel may have to wait for it (ranks sharing one GPU), and the peer for this rank
    const size_t np = (size_t)std::max<int64_t>(ctx->n, 1);
    DevBuf* four[] = {&ctx->b_kcount, &ctx->b_rdif, &ctx->b_thr_d2, &ctx->b_thr_idx};
    for (DevBuf* b : four)
      if (rc == CAB_OK && (b == &ctx->b_kcount || b == &ctx->b_rdif || max_nn_rsd > 0 || max_nn_normals > 0)) rc = reserve(ctx, *b, np * 4);
    if (rc == CAB_OK) rc = reserve(ctx, ctx->b_rsd, np * sizeof(float2));
    if (rc == CAB_OK) rc = reserve(ctx, ctx->b_stats, kStatBytes);
    if (rc == CAB_OK) rc = reserve(ctx, ctx->b_stats2, 65 * sizeof(float));
  }
  if (rc == CAB_OK && push) rc = comm_step_begin(ctx);
  if (rc == CAB_OK) rc = run_normals(ctx, (float)r, max_nn_normals, vp, nullptr, max_nn_normals == 0 ? max_nn_rsd : 0);
  if (rc == CAB_OK && halo) rc = comm_halo_send(ctx);
  if (rc == CAB_OK && push) rc = comm_step_before_push(ctx);
  if (halo) {
    // the halo rows' normals travel while the packets that do not read them run
    // (phase 1 runs on the compute stream; the halo's arrival and the boundary packets of phase 2 follow on the copy
    // stream, so that the boundary packets fill the SMs the interior kernel's last packets leave idle)
    if (rc == CAB_OK) rc = run_rsd(ctx, r, max_nn_rsd, ndiv, plane_radius, flags, 1);
    if (rc == CAB_OK && cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0) != cudaSuccess)
      rc = fail(ctx, CAB_ERR_CUDA, "cudaStreamWaitEvent failed");
    static const bool serial_phases = std::getenv("CAB_SERIAL_PHASES") != nullptr;  // A/B switch
    if (rc == CAB_OK) rc = comm_halo_receive(ctx, serial_phases ? ctx->stream : ctx->copy_stream);
    if (rc == CAB_OK) rc = run_rsd(ctx, r, max_nn_rsd, ndiv, plane_radius, flags, 2);
  } else if (rc == CAB_OK) {
    rc = run_rsd(ctx, r, max_nn_rsd, ndiv, plane_radius, flags);
  }
  if (rc == CAB_OK) rc = cudaEventRecord(ctx->ev[9], st) == cudaSuccess ? CAB_OK : fail(ctx, CAB_ERR_CUDA, "cudaEventRecord failed");
  if (rc == CAB_OK && push) rc = comm_step_end(ctx, plane_radius);
  if (rc == CAB_OK) rc = cudaEventRecord(ctx->ev[10], st) == cudaSuccess ? CAB_OK : fail(ctx, CAB_ERR_CUDA, "cudaEventRecord failed");
  // one synchronisation per step (also on the error path: nothing may be left running on the caller's buffers)
  if (rc != CAB_OK) cudaStreamSynchronize(ctx->copy_stream);  // a failed step may have left forked work unjoined
  const cudaError_t e = cudaStreamSynchronize(st);
  if (rc != CAB_OK) return rc;
  if (e != cudaSuccess) return fail(ctx, CAB_ERR_CUDA, "cab_step_normals_rsd: %s", cudaGetErrorString(e));
  if (ctx->slab)
    if (int rc2 = finish_slab(ctx)) return rc2;
  if (int rc2 = finish_pass_stats(ctx, 0)) return rc2;
  const int64_t k_normals = ctx->tm.neighbour_sum, c_normals = ctx->tm.candidate_sum;
  (void)k_normals;
  (void)c_normals;
  if (int rc2 = finish_pass_stats(ctx, 1)) return rc2;
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.exchange_ms, ctx->ev[9], ctx->ev[10]));
  CAB_CUDA(ctx, cudaEventElapsedTime(&ctx->tm.step_ms, ctx->ev[8], ctx->ev[10]));
  if (push) return comm_step_finish(ctx);
  return CAB_OK;
}

int cab_profile(const cab_ctx* ctx, cab_timings* out) {
  if (!ctx || !out) return CAB_ERR_ARG;
  *out = ctx->tm;
  return CAB_OK;
}

}  // extern "C"
