// plx_raygen.cuh — one ray from (camera pose, u, v): the arithmetic of generate_rays_batched, src/ray_sampling.py:212-264,
// shared by the stand-alone ray kernel (plx_eager.cu) and the fused training kernel (plx_train.cu).
#pragma once
#include "plx_device.cuh"

namespace plx {

// torch.linspace(0, 1, n)[i] as ATen's CPU kernel evaluates it (lower half step*i, upper half fma(-step, n-1-i, 1))
__device__ __forceinline__ float linspace01(int i, int n) {
    if (n <= 1) return 0.f;
    const float step = __fdiv_rn(1.f, (float)(n - 1));
    return i < n / 2 ? __fmul_rn(step, (float)i) : fmaf(-step, (float)(n - 1 - i), 1.f);
}

__device__ __forceinline__ float norm3_plain(float x, float y, float z) {     // strided pose columns: mul/add in order
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
__device__ __forceinline__ float norm3_fused(float x, float y, float z) {     // contiguous last axis: x^2 then two FMAs
    return __fsqrt_rn(fmaf(z, z, fmaf(y, y, __fmul_rn(x, x))));
}

struct RayOut {
    float dx, dy, dz;     // unit direction
    int up, vp;           // pixel column / row indices of the target lookup imgs[cam, vp, up]
};

// P = one (4,4) row-major camera-to-world matrix
__device__ __forceinline__ RayOut ray_from_uv(const float* __restrict__ P, float fov, float u, float v, int H, int W) {
    const float Xx = __ldg(P + 0), Xy = __ldg(P + 4), Xz = __ldg(P + 8);
    const float Yx = __ldg(P + 1), Yy = __ldg(P + 5), Yz = __ldg(P + 9);
    const float Zx = -__ldg(P + 2), Zy = -__ldg(P + 6), Zz = -__ldg(P + 10);
    const float aspect = __fdiv_rn(norm3_plain(Xx, Xy, Xz), norm3_plain(Yx, Yy, Yz));               // :218
    const float ua = __fmul_rn(fov, __fsub_rn(u, 0.5f));                                              // :234
    const float va = -__fmul_rn(__fmul_rn(fov, __fdiv_rn(1.f, aspect)), __fsub_rn(v, 0.5f));          // :235
    RayOut o;
    o.up = (int)fminf(rintf(__fmul_rn((float)H, u)), (float)(H - 1));                                 // :238
    o.vp = (int)fminf(rintf(__fmul_rn((float)W, v)), (float)(W - 1));                                 // :239
    const float dx = __fadd_rn(__fadd_rn(__fmul_rn(ua, Xx), __fmul_rn(va, Yx)), Zx);                  // :261
    const float dy = __fadd_rn(__fadd_rn(__fmul_rn(ua, Xy), __fmul_rn(va, Yy)), Zy);
    const float dz = __fadd_rn(__fadd_rn(__fmul_rn(ua, Xz), __fmul_rn(va, Yz)), Zz);
    const float nrm = norm3_fused(dx, dy, dz);
    o.dx = __fdiv_rn(dx, nrm);                                                                        // :262
    o.dy = __fdiv_rn(dy, nrm);
    o.dz = __fdiv_rn(dz, nrm);
    return o;
}

// target pixel `idx` (= (cam*H + v_pix)*W + u_pix) of the image set: fp32 RGBA as the reference keeps it, or the PNG's uint8 RGBA
// converted here exactly like src/data_processing.py:58 does it once for the whole set (fp32(u8) / 255, IEEE division)
__device__ __forceinline__ float4 load_target(const PlxRayGen& gen, int64_t idx) {
    if (gen.img_format == PLX_IMG_U8) {
        const uchar4 p = __ldg(reinterpret_cast<const uchar4*>(gen.imgs) + idx);
        return make_float4(__fdiv_rn((float)p.x, 255.f), __fdiv_rn((float)p.y, 255.f), __fdiv_rn((float)p.z, 255.f),
                           __fdiv_rn((float)p.w, 255.f));
    }
    return __ldg(reinterpret_cast<const float4*>(gen.imgs) + idx);
}

}  // namespace plx
