// fake_kernels.cpp — model kernels for the host-pipeline tests.  TEST INFRASTRUCTURE ONLY (see fake_cuda.h).
//
// The launch entry points of pl-svo_b200/csrc/internal.h, implemented on the host: each "kernel" is queued on its stream
// in the model runtime and, when it runs, checks that every byte the real kernel would read lies inside a device block
// and folds those bytes into a digest per work item.  The digest comes back through the ordinary output arrays, so a test
// sees — through the unchanged C ABI and Python mirror — exactly which bytes the host code presented to the kernel:
//
//   alignment : out_n_tracked[b] = digest of pair b, as (h >> 1) | 1 (image levels min..max of both frames, poses, counts, every feature
//               array that was shipped); out_H[b][2*(l-min_level)+{0,1}] = digest of the ref / cur level l alone (top 52
//               bits, exact in a double); out_T[b] = T_cur_w[b]; out_patch_iters/levels = the pair's feature counts.
//               Levels that the real kernel derives itself (AlignArgs::derive_from) are derived here the same way; the
//               arrival gate is honoured pair by pair, advancing the other streams only as far as the gate requires.
//   pyramid   : the real computation (truncating 2x2 mean), it is the producer of derived levels.
//   pose-opt  : out_num_pt[b] = digest of frame b's inputs; out_T[b] = T_f_w[b].
//   the next-row kernels report cudaErrorNotSupported (their host code is plain upload -> launch -> download).
//
// The same digest is restated in NumPy in tests/test_host_pipeline_cpu.py, from the caller's arrays.
#include <string.h>

#include <memory>

#include "../../pl-svo_b200/csrc/internal.h"
#include "fake_cuda.h"

namespace {

constexpr uint64_t K = 0x9E3779B97F4A7C15ull;

// position- and value-sensitive, vectorisable: sum over i of (byte_i + 1) * ((i + 1) * K)   (mod 2^64)
uint64_t dig_bytes(const void* p, size_t n, uint64_t first_index = 0) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  uint64_t h = 0;
  for (size_t i = 0; i < n; ++i) h += (uint64_t)(b[i] + 1u) * ((first_index + i + 1) * K);
  return h;
}

uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + K + (h << 6) + (h >> 2);
  return h;
}

// rows x cols bytes of a pitched image, indexed row-major over the visible pixels only
uint64_t dig_image(const uint8_t* img, int rows, int cols, size_t pitch) {
  uint64_t h = 0;
  for (int y = 0; y < rows; ++y) h += dig_bytes(img + (size_t)y * pitch, (size_t)cols, (uint64_t)y * cols);
  return h;
}

void half_sample(const uint8_t* src, size_t pin, uint8_t* dst, size_t pout, int rows, int cols) {
  for (int y = 0; y < rows; ++y) {
    const uint8_t* r0 = src + (size_t)(2 * y) * pin;
    for (int x = 0; x < cols; ++x)
      dst[(size_t)y * pout + x] = (uint8_t)(((int)r0[2 * x] + (int)r0[2 * x + 1] + (int)r0[pin + 2 * x] + (int)r0[pin + 2 * x + 1]) >> 2);
  }
}

template <class T>
uint64_t dig_array(uint64_t h, const T* base, size_t first, size_t count, const char* what) {
  if (!base) return mix(h, 0x5151);  // array not shipped
  const T* p = base + first;
  if (!fakecuda::check(p, count * sizeof(T), what)) return mix(h, 0xDEAD);
  return mix(h, dig_bytes(p, count * sizeof(T)));
}

struct AlignRun {
  plsvo::AlignArgs a;
  cudaStream_t stream;
  int next = 0;
};

bool align_pair(const plsvo::AlignArgs& a, int b) {
  using fakecuda::check;
  bool ok = true;
  // levels the kernel forms itself from the finest shipped one (gated host pipeline)
  if (a.derive_from >= 0) {
    for (int l = a.derive_from + 1; l <= a.max_level; ++l) {
      const int cols = a.width >> l, rows = a.height >> l;
      for (int which = 0; which < 2; ++which) {
        const uint8_t* src = (which ? a.cur_img[l - 1] : a.ref_img[l - 1]) + (size_t)b * a.stride[l - 1];
        uint8_t* dst = const_cast<uint8_t*>(which ? a.cur_img[l] : a.ref_img[l]) + (size_t)b * a.stride[l];
        if (!check(src, a.stride[l - 1], "align kernel: source level of a derived level") ||
            !check(dst, a.stride[l], "align kernel: derived level"))
          return false;
        half_sample(src, a.pitch[l - 1], dst, a.pitch[l], rows, cols);
      }
    }
  }
  uint64_t h = 0;
  for (int i = 0; i < 36; ++i) a.out_H[(size_t)b * 36 + i] = 0.0;
  for (int l = a.min_level; l <= a.max_level; ++l) {
    const int cols = a.width >> l, rows = a.height >> l;
    if (!a.ref_img[l] || !a.cur_img[l] || !a.pitch[l]) {
      fakecuda::error("align kernel: a level in [min_level, max_level] has no image");
      return false;
    }
    if (a.pitch[l] % 4 != 0 || a.stride[l] % 16 != 0) fakecuda::error("align kernel: level pitch / stride not word / 16-byte aligned");
    const uint8_t* r = a.ref_img[l] + (size_t)b * a.stride[l];
    const uint8_t* c = a.cur_img[l] + (size_t)b * a.stride[l];
    ok &= check(r, a.stride[l], "align kernel: reference image level");
    ok &= check(c, a.stride[l], "align kernel: current image level");
    if (!ok) return false;
    const uint64_t hr = dig_image(r, rows, cols, a.pitch[l]), hc = dig_image(c, rows, cols, a.pitch[l]);
    h = mix(mix(h, hr), hc);
    const int k = 2 * (l - a.min_level);
    if (k + 1 < 36) a.out_H[(size_t)b * 36 + k] = (double)(hr >> 12), a.out_H[(size_t)b * 36 + k + 1] = (double)(hc >> 12);
  }
  const size_t np_ = (size_t)a.n_pts, ns_ = (size_t)a.n_segs;
  const int np = a.pt_count ? a.pt_count[b] : a.n_pts, ns = a.seg_count ? a.seg_count[b] : a.n_segs;
  if (np < 0 || np > a.n_pts || ns < 0 || ns > a.n_segs) {
    fakecuda::error("align kernel: feature count outside [0, n]");
    return false;
  }
  h = dig_array(h, a.T_ref_w, (size_t)b * 7, 7, "T_ref_w");
  h = dig_array(h, a.T_cur_w, (size_t)b * 7, 7, "T_cur_w");
  h = mix(h, (uint64_t)np * 65536u + (uint64_t)ns);
  h = dig_array(h, a.pt_px, b * np_ * 2, (size_t)np * 2, "pt_px");
  h = dig_array(h, a.pt_f, b * np_ * 3, (size_t)np * 3, "pt_f");
  h = dig_array(h, a.pt_pos, b * np_ * 3, (size_t)np * 3, "pt_pos");
  h = dig_array(h, a.pt_depth, b * np_, (size_t)np, "pt_depth");
  h = dig_array(h, a.pt_valid, b * np_, (size_t)np, "pt_valid");
  h = dig_array(h, a.seg_spx, b * ns_ * 2, (size_t)ns * 2, "seg_spx");
  h = dig_array(h, a.seg_epx, b * ns_ * 2, (size_t)ns * 2, "seg_epx");
  h = dig_array(h, a.seg_sf, b * ns_ * 3, (size_t)ns * 3, "seg_sf");
  h = dig_array(h, a.seg_ef, b * ns_ * 3, (size_t)ns * 3, "seg_ef");
  h = dig_array(h, a.seg_spos, b * ns_ * 3, (size_t)ns * 3, "seg_spos");
  h = dig_array(h, a.seg_epos, b * ns_ * 3, (size_t)ns * 3, "seg_epos");
  h = dig_array(h, a.seg_sdepth, b * ns_, (size_t)ns, "seg_sdepth");
  h = dig_array(h, a.seg_edepth, b * ns_, (size_t)ns, "seg_edepth");
  h = dig_array(h, a.seg_length, b * ns_, (size_t)ns, "seg_length");
  h = dig_array(h, a.seg_valid, b * ns_, (size_t)ns, "seg_valid");
  // outputs, every one of them range-checked like an input
  ok &= check(a.out_T + (size_t)b * 7, 56, "out_T") && check(a.out_n_tracked + b, 8, "out_n_tracked") &&
        check(a.out_H + (size_t)b * 36, 288, "out_H") && check(a.out_iters + (size_t)b * PLSVO_MAX_LEVELS, 4 * PLSVO_MAX_LEVELS, "out_iters") &&
        check(a.out_status + b, 4, "out_status") && check(a.out_patch_iters + b, 4, "out_patch_iters") &&
        check(a.out_patch_levels + b, 4, "out_patch_levels");
  if (a.n_segs > 0) ok &= check(a.out_seg_killed + b * ns_, ns_, "out_seg_killed");
  if (!ok) return false;
  for (int i = 0; i < 7; ++i) a.out_T[(size_t)b * 7 + i] = a.T_cur_w ? a.T_cur_w[(size_t)b * 7 + i] : 0.0;
  a.out_n_tracked[b] = (long long)((h >> 1) | 1u);  // positive, like a count of tracked patches
  for (int l = 0; l < PLSVO_MAX_LEVELS; ++l) a.out_iters[(size_t)b * PLSVO_MAX_LEVELS + l] = (l >= a.min_level && l <= a.max_level) ? 1 : 0;
  a.out_status[b] = 0;
  a.out_patch_iters[b] = (uint32_t)np;
  a.out_patch_levels[b] = (uint32_t)ns;
  for (int j = 0; j < a.n_segs; ++j) a.out_seg_killed[b * ns_ + j] = 0;
  return true;
}

}  // namespace

namespace plsvo {

// Same structure as make_layout() in align_kernel.cu (which is device code and cannot be compiled here); only the plan's
// decisions depend on it (variant choice, image staging), none of which the digest kernel looks at.
size_t align_smem_bytes(int n_pts, int n_segs, int max_patches, int max_seg_slots, int img_bytes, int threads) {
  const size_t nw = (size_t)threads / 32, chunks = ((size_t)n_pts + 31) / 32, rounds = ((size_t)n_pts + threads - 1) / threads;
  size_t o = 1024;                                   // control block
  o += nw * 256 + 256 + 8 * (rounds * nw + 1);       // reduction scratch, chunk totals
  o += 256 * (chunks + 1) + 4 * (chunks + 1) + 512;  // chi2 items
  o += 64 * 48;                                      // opaque patches
  o += 21 * (size_t)n_segs + 2 * (size_t)max_seg_slots + (size_t)n_pts;
  o += 24 * (size_t)max_patches + 64 * (size_t)threads;
  o = (o + 127) / 128 * 128;
  return o + (size_t)img_bytes + 16;
}

cudaError_t align_kernel_prepare(int threads, int min_blocks, size_t smem_bytes, int* ctas_per_sm) {
  (void)threads;
  *ctas_per_sm = (int)std::min<size_t>((size_t)min_blocks, (232448 + 1024) / (smem_bytes + 1024));
  return cudaSuccess;
}

cudaError_t weight_selftest_launch(uint32_t, uint32_t, unsigned long long* d_mismatch, cudaStream_t s) {
  return fakecuda::enqueue(s, [d_mismatch]() {
    if (fakecuda::check(d_mismatch, 8, "selftest counter")) *d_mismatch = 0;
    return true;
  });
}

cudaError_t align_kernel_launch(const AlignArgs& a, int grid, int threads, int min_blocks, size_t smem_bytes, cudaStream_t s) {
  if (grid < 1 || threads < 32 || min_blocks < 1 || smem_bytes > 232448) {
    fakecuda::error("align_kernel_launch: launch configuration out of range");
    return cudaErrorInvalidConfiguration;
  }
  auto run = std::make_shared<AlignRun>();
  run->a = a, run->stream = s;
  return fakecuda::enqueue(s, [run]() {
    const AlignArgs& a = run->a;
    if (run->next == 0) {
      if (!fakecuda::check(a.work_counter, 4, "work counter")) return true;
      if (*a.work_counter != 0) fakecuda::error("align kernel: work counter not cleared before the launch");
    }
    while (run->next < a.B) {
      const int b = run->next;
      if (a.gate_chunk > 0) {
        if (!fakecuda::check(a.arrived, 4, "arrival counter")) return true;
        const unsigned need = (unsigned)(b / a.gate_chunk) + 1u;
        while (*a.arrived < need)
          if (!fakecuda::advance_others(run->stream)) return false;  // blocked: this pair's chunk is still in flight
      }
      align_pair(a, b);
      run->next = b + 1;
      *a.work_counter = (unsigned)run->next;
    }
    return true;
  });
}

size_t poseopt_smem_bytes(int n_pts, int n_segs) { return 4096 + 24 * ((size_t)n_pts + (size_t)n_segs); }

cudaError_t poseopt_kernel_launch(const PoseOptArgs& a0, size_t, cudaStream_t s) {
  const PoseOptArgs a = a0;
  return fakecuda::enqueue(s, [a]() {
    const size_t np_ = (size_t)a.n_pts, ns_ = (size_t)a.n_segs;
    for (int b = 0; b < a.B; ++b) {
      const int np = a.pt_count ? a.pt_count[b] : a.n_pts, ns = a.seg_count ? a.seg_count[b] : a.n_segs;
      uint64_t h = dig_array<double>(0, a.T_f_w, (size_t)b * 7, 7, "T_f_w");
      h = mix(h, (uint64_t)np * 65536u + (uint64_t)ns);
      h = dig_array(h, a.pt_f, b * np_ * 3, (size_t)np * 3, "pt_f");
      h = dig_array(h, a.pt_pos, b * np_ * 3, (size_t)np * 3, "pt_pos");
      h = dig_array(h, a.pt_level, b * np_, (size_t)np, "pt_level");
      h = dig_array(h, a.pt_valid, b * np_, (size_t)np, "pt_valid");
      h = dig_array(h, a.seg_line, b * ns_ * 3, (size_t)ns * 3, "seg_line");
      h = dig_array(h, a.seg_spos, b * ns_ * 3, (size_t)ns * 3, "seg_spos");
      h = dig_array(h, a.seg_epos, b * ns_ * 3, (size_t)ns * 3, "seg_epos");
      h = dig_array(h, a.seg_level, b * ns_, (size_t)ns, "seg_level");
      h = dig_array(h, a.seg_valid, b * ns_, (size_t)ns, "seg_valid");
      if (!fakecuda::check(a.out_T + (size_t)b * 7, 56, "poseopt out_T") || !fakecuda::check(a.out_num_pt + b, 8, "poseopt out_num_pt") ||
          !fakecuda::check(a.out_status + b, 4, "poseopt out_status"))
        return true;
      for (int i = 0; i < 7; ++i) a.out_T[(size_t)b * 7 + i] = a.T_f_w[(size_t)b * 7 + i];
      a.out_num_pt[b] = (long long)((h >> 1) | 1u);
      a.out_status[b] = 0;
    }
    return true;
  });
}

cudaError_t pyramid_kernel_launch(const PyramidArgs& a0, cudaStream_t s) {
  const PyramidArgs a = a0;
  return fakecuda::enqueue(s, [a]() {
    for (int b = 0; b < a.B; ++b)
      for (int l = 1; l < a.n_levels; ++l) {
        const int cols = a.width >> l, rows = a.height >> l;
        const uint8_t* src = a.level[l - 1] + (size_t)b * a.stride[l - 1];
        uint8_t* dst = a.level[l] + (size_t)b * a.stride[l];
        const size_t sspan = (size_t)(2 * rows - 1) * a.pitch[l - 1] + 2 * (size_t)cols;
        const size_t dspan = (size_t)(rows - 1) * a.pitch[l] + (size_t)cols;
        if (rows <= 0 || cols <= 0) continue;
        if (!fakecuda::check(src, sspan, "pyramid kernel: source level") || !fakecuda::check(dst, dspan, "pyramid kernel: destination level"))
          return true;
        half_sample(src, a.pitch[l - 1], dst, a.pitch[l], rows, cols);
      }
    return true;
  });
}

cudaError_t align2d_kernel_launch(const Align2DArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t align1d_kernel_launch(const Align2DArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t match_direct_kernel_launch(const MatchArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t seed_update_kernel_launch(const SeedArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t line_seed_update_kernel_launch(const SeedArgs&, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t structopt_kernel_launch(const StructOptArgs&, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace plsvo
