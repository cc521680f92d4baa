// Probe: is the texture unit's bilinear filter on UNORM8x4 texels exact enough to recover the
// cv2.remap integer blend  (sum_i w_i p_i + 512) >> 10  with 5-bit fractions?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hwfilter_probe tools/hwfilter_probe.cu
// Prints the number of (texel position, fx, fy, channel) cases whose recovered integer sum differs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void probe(cudaTextureObject_t tex, const uchar4 *img, int W, int H, unsigned long long *bad,
                      unsigned long long *bad_out, float *maxerr) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;  // texel position
    const int y = blockIdx.y;
    if (x >= W - 1 || y >= H - 1) return;
    const uchar4 p00 = img[y * W + x], p01 = img[y * W + x + 1], p10 = img[(y + 1) * W + x], p11 = img[(y + 1) * W + x + 1];
    unsigned long long nb = 0, nbo = 0;
    float me = 0.f;
    for (int fy = 0; fy < 32; ++fy)
        for (int fx = 0; fx < 32; ++fx) {
            const float u = (float)x + 0.5f + fx * 0.03125f, v = (float)y + 0.5f + fy * 0.03125f;
            const float4 r = tex2D<float4>(tex, u, v);
            const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
            const float rr[3] = {r.x, r.y, r.z};
            const int a[3] = {p00.x, p00.y, p00.z}, b[3] = {p01.x, p01.y, p01.z}, c[3] = {p10.x, p10.y, p10.z},
                      d[3] = {p11.x, p11.y, p11.z};
            for (int ch = 0; ch < 3; ++ch) {
                const int S = a[ch] * w00 + b[ch] * w01 + c[ch] * w10 + d[ch] * w11;
                const float est = rr[ch] * 261120.0f;  // 255 * 1024
                const int Sr = __float2int_rn(est);
                me = fmaxf(me, fabsf(est - (float)S));
                nb += (Sr != S);
                nbo += (((Sr + 512) >> 10) != ((S + 512) >> 10));
            }
        }
    if (nb) atomicAdd(bad, nb);
    if (nbo) atomicAdd(bad_out, nbo);
    atomicMax((int *)maxerr, __float_as_int(me));
}

int main() {
    const int W = 512, H = 256;
    uchar4 *h = (uchar4 *)malloc(W * H * 4);
    srand(1);
    for (int i = 0; i < W * H; ++i) h[i] = make_uchar4(rand() & 255, rand() & 255, rand() & 255, 0);
    for (int i = 0; i < 64; ++i) h[i] = make_uchar4(i & 1 ? 255 : 0, 255, i & 2 ? 255 : 254, 0);  // extremes
    cudaArray_t arr;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    cudaMallocArray(&arr, &fd, W, H);
    cudaMemcpy2DToArray(arr, 0, 0, h, W * 4, W * 4, H, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    uchar4 *d_img;
    cudaMalloc(&d_img, W * H * 4);
    cudaMemcpy(d_img, h, W * H * 4, cudaMemcpyHostToDevice);
    unsigned long long *d_bad;
    float *d_me;
    cudaMalloc(&d_bad, 16);
    cudaMalloc(&d_me, 4);
    cudaMemset(d_bad, 0, 16);
    cudaMemset(d_me, 0, 4);
    probe<<<dim3((W + 127) / 128, H), 128>>>(tex, d_img, W, H, d_bad, d_bad + 1, d_me);
    unsigned long long bad[2];
    float me;
    cudaMemcpy(bad, d_bad, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&me, d_me, 4, cudaMemcpyDeviceToHost);
    const unsigned long long total = (unsigned long long)(W - 1) * (H - 1) * 1024 * 3;
    printf("{\"cases\": %llu, \"sum_mismatch\": %llu, \"output_mismatch\": %llu, \"max_abs_err_in_sum_units\": %g, \"cuda_error\": \"%s\"}\n",
           total, bad[0], bad[1], me, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
