// Training-batch preparation on the GPU: get_data_from_chunk_v2, myTool.py:1158-1199 -- per image
// RandomResizeLong (cv2.resize, bilinear, :995-1008) -> flip (:895-899) -> ImageNet normalisation (:1181-1183) -> RandomCrop
// into a zero-filled crop_size x crop_size container (:923-955) -- as ONE gather kernel over the output pixels: the
// resized / flipped / cropped intermediate images never exist, every output pixel maps back to four source pixels of the
// decoded uint8 image.  The random draws stay on the host (acr_wsss_b200/data.py mirrors the reference's RNG call order)
// and arrive here as per-image parameters.  HBM-bound: reads <= 4 source bytes x 3 per output pixel (L2-resident), writes
// 12 (+3) bytes per output pixel.
#include "common.cuh"

namespace {

struct ImgParam {          // one per image, 12 ints (see include/acr_b200.h)
  int h, w;                // decoded image
  int th, tw;              // size after RandomResizeLong
  int flip;                // np.fliplr of the resized image
  int img_top, img_left;   // crop origin inside the resized image
  int cont_top, cont_left; // paste origin inside the container
  int ch, cw;              // pasted extent
  int pad_;
};

// cv2.resize INTER_LINEAR source coordinate (float path): fx = (dx + 0.5) * scale - 0.5, floor, clamp with zero weight
__device__ __forceinline__ void src_coord(int d, double scale, int n, int& s0, int& s1, float& f) {
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(fx);
  fx -= (float)s;
  if (s < 0) { s = 0; fx = 0.f; }
  if (s >= n - 1) { s = n - 1; fx = 0.f; }
  s0 = s;
  s1 = min(s + 1, n - 1);
  f = fx;
}

__global__ void __launch_bounds__(256)
augment_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ offsets, const ImgParam* __restrict__ params,
               int dim, float* __restrict__ out, unsigned char* __restrict__ ori) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dim || y >= dim) return;
  const ImgParam p = params[b];
  const size_t plane = (size_t)dim * dim;
  float* o = out + (size_t)b * 3 * plane + (size_t)y * dim + x;
  unsigned char* oo = ori ? ori + (size_t)b * 3 * plane + (size_t)y * dim + x : nullptr;
  const int cy = y - p.cont_top, cx = x - p.cont_left;
  if (cy < 0 || cy >= p.ch || cx < 0 || cx >= p.cw) {        // outside the pasted region: the container's zeros
    o[0] = 0.f; o[plane] = 0.f; o[2 * plane] = 0.f;
    if (oo) {   // the reference de-normalises the zero-filled container: (0*std + mean)*255, truncated (myTool.py:1188-1192)
      oo[0] = (unsigned char)(0.485f * 255.f); oo[plane] = (unsigned char)(0.456f * 255.f); oo[2 * plane] = (unsigned char)(0.406f * 255.f);
    }
    return;
  }
  const int ry = p.img_top + cy;
  int rx = p.img_left + cx;
  if (p.flip) rx = p.tw - 1 - rx;
  int y0, y1, x0, x1;
  float fy, fx;
  src_coord(ry, (double)p.h / (double)p.th, p.h, y0, y1, fy);
  src_coord(rx, (double)p.w / (double)p.tw, p.w, x0, x1, fx);
  const unsigned char* im = src + offsets[b];
  const unsigned char* r0 = im + (size_t)y0 * p.w * 3;
  const unsigned char* r1 = im + (size_t)y1 * p.w * 3;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // horizontal pass on both rows, then vertical (cv2's order)
    const float a = (float)r0[x0 * 3 + c] * (1.f - fx) + (float)r0[x1 * 3 + c] * fx;
    const float bb = (float)r1[x0 * 3 + c] * (1.f - fx) + (float)r1[x1 * 3 + c] * fx;
    const float v = a * (1.f - fy) + bb * fy;
    const float nv = (v / 255.f - mean[c]) / stdv[c];
    o[c * plane] = nv;
    if (oo) oo[c * plane] = (unsigned char)fminf(fmaxf((nv * stdv[c] + mean[c]) * 255.f, 0.f), 255.f);
  }
}

}  // namespace

extern "C" int acr_augment_batch(const unsigned char* src, const long long* offsets, const int* params, int B, int crop_size,
                                 float* out, unsigned char* ori_out, void* stream) {
  ACR_REQUIRE(src && offsets && params && out, ACR_E_INVAL, "acr_augment_batch: null pointer");
  ACR_REQUIRE(B > 0 && B <= 65535 && crop_size > 0, ACR_E_INVAL, "acr_augment_batch: bad shape");
  static_assert(sizeof(ImgParam) == 12 * sizeof(int), "parameter record is 12 ints");
  dim3 grid((crop_size + 31) / 32, (crop_size + 7) / 8, B);
  augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, offsets, reinterpret_cast<const ImgParam*>(params), crop_size, out, ori_out);
  return acr::check_launch("augment_kernel");
}
