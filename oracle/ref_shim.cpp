// extern "C" door onto the UNMODIFIED reference C++ (compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/).  Test infrastructure only; contains no reference code.
#include "bilateralfilter.hpp"

extern "C" void ref_bilateralfilter_batch(float* images, int len_images, float* ins, int len_ins, float* outs, int len_outs,
                                          int N, int K, int H, int W, float sigmargb, float sigmaxy) {
  bilateralfilter_batch(images, len_images, ins, len_ins, outs, len_outs, N, K, H, W, sigmargb, sigmaxy);
}
