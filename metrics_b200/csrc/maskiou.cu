// K12 — instance-mask IoU for `MeanAveragePrecision(iou_type="segm")` (reference seam: detection/mean_ap.py:848-853 turns every
// mask into a run-length code on the HOST with pycocotools' `mask_utils.encode`, `COCOeval.computeIoU` then walks pairs of
// run-length codes on the host again, maskApi.c:rleIou).
//
// Here masks never leave the device: `mb200_mask_pack_bits` turns a batch of boolean masks into one bit per pixel (32 pixels
// per word, ballot-packed) and counts their areas; `mb200_mask_pair_intersections` produces, for every image, the
// [detections x ground truths] table of intersection pixel counts = popcount(a & b) over the words of the two masks — integer
// exact, HBM/L2-bound (each pair reads its two bit rows once; 32 pixels per 8 bytes moved).  The matching kernel
// (cocomap.cu, `mb200_coco_map_match_ex`) derives the IoU from the table and the areas exactly like maskApi.c:rleIou does
// (intersection 0 -> 0; crowd ground truth -> union = detection area).
#include "common.cuh"

namespace mb200 {

extern void count_launch();

// one warp per 32 pixels per step: lane l reads pixel 32 w + l, the ballot is word w
__global__ void __launch_bounds__(256) mask_pack_bits_kernel(const unsigned char* __restrict__ masks, long long n_masks,
                                                             long long hw, long long words, unsigned* __restrict__ out,
                                                             long long out_stride, long long* __restrict__ area) {
    const int lane = threadIdx.x & 31;
    const long long warps_per_grid = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total_words = n_masks * words;
    unsigned long long local_area = 0;
    long long last_mask = -1;
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total_words; w += warps_per_grid) {
        const long long m = w / words, ww = w % words;
        const long long px = ww * 32 + lane;
        const bool bit = px < hw && masks[m * hw + px] != 0;
        const unsigned word = __ballot_sync(kFull, bit);
        if (lane == 0) {
            out[m * out_stride + ww] = word;
            if (m != last_mask) {
                if (last_mask >= 0 && local_area) atomicAdd(reinterpret_cast<unsigned long long*>(area + last_mask), local_area);
                last_mask = m;
                local_area = 0;
            }
            local_area += (unsigned)__popc(word);
        }
    }
    if (lane == 0 && last_mask >= 0 && local_area) atomicAdd(reinterpret_cast<unsigned long long*>(area + last_mask), local_area);
}

// grid = (images, splits): the warps of the CTAs of one image share its detection x ground-truth pairs; a pair of different
// classes (unless micro) is never looked at by the matcher and stays 0.
__global__ void __launch_bounds__(256) mask_pair_inter_kernel(const unsigned* __restrict__ det_words,
                                                              const long long* __restrict__ det_word_off,
                                                              const unsigned* __restrict__ gt_words,
                                                              const long long* __restrict__ gt_word_off,
                                                              const int* __restrict__ det_off, const int* __restrict__ gt_off,
                                                              const int* __restrict__ img_words,
                                                              const long long* __restrict__ det_label,
                                                              const long long* __restrict__ gt_label, int micro,
                                                              const long long* __restrict__ pair_off,
                                                              double* __restrict__ inter) {
    const int img = blockIdx.x;
    const int d0 = det_off[img], D = det_off[img + 1] - d0;
    const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
    const int words = img_words[img];
    const long long pairs = (long long)D * G;
    const int lane = threadIdx.x & 31;
    const long long wstep = (long long)gridDim.y * (blockDim.x >> 5);
    for (long long pr = (long long)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); pr < pairs; pr += wstep) {
        const int d = (int)(pr / G), g = (int)(pr % G);
        unsigned cnt = 0;
        if (micro || det_label[d0 + d] == gt_label[g0 + g]) {
            const unsigned* __restrict__ a = det_words + det_word_off[d0 + d];
            const unsigned* __restrict__ b = gt_words + gt_word_off[g0 + g];
            for (int w = lane; w < words; w += 32) cnt += (unsigned)__popc(a[w] & b[w]);
            cnt = __reduce_add_sync(kFull, cnt);
        }
        if (lane == 0) inter[pair_off[img] + pr] = (double)cnt;
    }
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_mask_pack_bits(const uint8_t* masks, int64_t n_masks, int64_t pixels_per_mask, uint32_t* words_out,
                                    int64_t out_stride_words, int64_t* area_out, void* stream) {
    MB200_REQUIRE(n_masks >= 0 && pixels_per_mask >= 0, "bad sizes");
    const int64_t words = (pixels_per_mask + 31) / 32;
    MB200_REQUIRE(out_stride_words >= words, "output row stride %lld is smaller than the %lld words of a mask",
                  (long long)out_stride_words, (long long)words);
    if (n_masks == 0) return 0;
    MB200_REQUIRE(area_out, "NULL pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MB200_CUDA_OK(cudaMemsetAsync(area_out, 0, (size_t)n_masks * 8, st));
    if (words == 0) return 0;
    MB200_REQUIRE(masks && words_out, "NULL pointer");
    long long grid = (n_masks * words + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    mask_pack_bits_kernel<<<(unsigned)grid, 256, 0, st>>>(masks, n_masks, pixels_per_mask, words, words_out, out_stride_words,
                                                         reinterpret_cast<long long*>(area_out));
    count_launch();
    return check_cuda(cudaGetLastError(), "mask pack launch");
}

extern "C" int mb200_mask_pair_intersections(const uint32_t* det_words, const int64_t* det_word_off, const uint32_t* gt_words,
                                             const int64_t* gt_word_off, const int32_t* det_off, const int32_t* gt_off,
                                             const int32_t* img_words, const int64_t* det_label, const int64_t* gt_label,
                                             int micro, const int64_t* pair_off, int64_t n_img, int64_t max_pairs_per_img,
                                             double* inter_out, void* stream) {
    MB200_REQUIRE(n_img >= 0 && max_pairs_per_img >= 0, "bad sizes");
    if (n_img == 0 || max_pairs_per_img == 0) return 0;
    MB200_REQUIRE(det_word_off && gt_word_off && det_off && gt_off && img_words && det_label && gt_label && pair_off && inter_out,
                  "NULL pointer");
    // enough CTAs per image that the busiest image's pairs are spread, without flooding the grid for thousands of images
    long long splits = (max_pairs_per_img + 7) / 8;
    const long long want = ((long long)sm_count() * 8 + n_img - 1) / n_img;
    if (splits > want) splits = want;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    mask_pair_inter_kernel<<<dim3((unsigned)n_img, (unsigned)splits), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        det_words, reinterpret_cast<const long long*>(det_word_off), gt_words, reinterpret_cast<const long long*>(gt_word_off),
        det_off, gt_off, img_words, reinterpret_cast<const long long*>(det_label), reinterpret_cast<const long long*>(gt_label),
        micro, reinterpret_cast<const long long*>(pair_off), inter_out);
    count_launch();
    return check_cuda(cudaGetLastError(), "mask pair launch");
}
