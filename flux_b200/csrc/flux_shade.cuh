// flux_shade.cuh — BRDF sampling with the reference's operation order.
//
//   Lambertian::sample_f       fluxcore/src/brdf.rs:20-30
//   PerfectSpecular::sample_f  fluxcore/src/brdf.rs:39-45
//   GlossySpecular::sample_f   fluxcore/src/brdf.rs:55-78
//   to_unit_hemi               samplers/src/lib.rs:133-142
//
// Each returns the child direction wi and the weight the material applies to
// the child's radiance: result = (f (*) L_child) * (n.wi / pdf)
// (materials.rs:31-32, 69-70).  f is returned by reference.
#pragma once
#include "flux_scene.cuh"

// Matte: materials.rs:19-33 + brdf.rs:20-30.  f is the per-material constant
// (cd*kd)*INV_PI precomputed on the host (same IEEE products).
__device__ __forceinline__ void matte_sample(V3 normal, V3 hemi, V3 &wi, double &weight) {
    V3 w = normal;
    V3 v = normalize3_dev(cross3(mk3(0.0034, 1.0, 0.0071), w));
    V3 u = cross3(v, w);
    wi = normalize3_dev((hemi.x * u + hemi.y * v) + hemi.z * w);
    double pdf = dot3(normal, wi) * FLUX_INV_PI;
    double ndotwi = dot3(normal, wi);
    weight = div_full(ndotwi, pdf);
}

// Reflective + PerfectSpecular: materials.rs:57-71 + brdf.rs:39-45.
__device__ __forceinline__ void specular_sample(V3 normal, V3 dir, V3 &wi, double &weight) {
    V3 wo = dir * -1.0;
    double ndotwo = dot3(normal, wo);
    wi = neg3(wo) + normal * ndotwo * 2.0;
    double pdf = dot3(normal, wi);
    weight = div_full(dot3(normal, wi), pdf);
}

// samplers/src/lib.rs:133-142 with inv_e1 = 1.0/(e+1.0) precomputed.
__device__ __forceinline__ V3 to_unit_hemi_dev(double px, double py, double inv_e1) {
    double sin_phi, cos_phi;
    sincos((2.0 * FLUX_PI) * px, &sin_phi, &cos_phi);
    double cos_theta = pow(1.0 - py, inv_e1);
    double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    double pu = sin_theta * cos_phi;
    double pv = sin_theta * sin_phi;
    double pw = cos_theta;
    return normalize3_dev(mk3(pu, pv, pw));
}

// x^y for the glossy lobe (brdf.rs:73, `powf`).  The reference's value comes from the platform libm and CUDA's pow
// differs from it in the last ulp anyway (glossy radiance is pinned at 1e-6 relative, not bit-exact), so for the
// ordinary case 0 < x < inf the lobe is exp(y * log(x)): |y ln x| stays below ~40 on the path (larger values
// underflow the lobe), so the relative error is |y ln x| * 2^-52 < 1e-14 — eight orders inside the bar — for a
// third of pow()'s instructions (pow was the critical path of the sorted shading stage).  Zero,
// negative, infinite and NaN bases keep pow()'s special-case semantics.
__device__ __forceinline__ double lobe_pow(double x, double y) {
    if (x > 0.0 && x < 1.0e300) return exp(y * log(x));
    return pow(x, y);
}

// Reflective + GlossySpecular: materials.rs:57-71 + brdf.rs:55-78, given hs = to_unit_hemi(pixel_sample, exp)
// (brdf.rs:64).  lobe multiplies the per-material constant cs*ks to give f.
__device__ __forceinline__ void glossy_sample_hs(V3 normal, V3 dir, V3 hs, double ex, V3 &wi, double &weight,
                                                 double &lobe, bool &flipped) {
    V3 wo = dir * -1.0;
    double ndotwo = dot3(normal, wo);
    V3 r = neg3(wo) + normal * ndotwo * 2.0;
    V3 w = r;
    V3 u = normalize3_dev(cross3(mk3(0.00424, 1.0, 0.00764), w));
    V3 v = cross3(u, w);
    V3 wi0 = (u * hs.x + v * hs.y) + w * hs.z;
    flipped = dot3(normal, wi0) < 0.0;
    if (flipped)
        wi = (u * -hs.x - v * hs.y) + w * hs.z;
    else
        wi = wi0;
    const double rdotwi = dot3(r, wi);
    const double ndotwi = dot3(normal, wi);
#ifndef FLUX_GLOSSY_EXACT_LOBE
    // The lobe value x^e enters the result twice, as f = (cs*ks)*lobe and through pdf = lobe*ndotwi in the weight
    // ndotwi/pdf (brdf.rs:73-77, materials.rs:69-70): ((cs*ks*lobe) (*) L) * (ndotwi / (lobe*ndotwi)).  Whenever the
    // lobe is an ordinary double — finite, non-zero, far from the denormal range — that product equals
    // ((cs*ks) (*) L) * (ndotwi/ndotwi) up to five roundings (6e-16 relative), nine orders inside the 1e-6 bar that
    // applies to glossy radiance, so the power function is not evaluated at all: lobe := 1.  A cheap f32 guard
    // (|e * log2 x| < 900, x > 0, 1e-200 < |ndotwi| < 1e200) decides; everything else — zero / negative / non-finite
    // bases, lobes that underflow (their 0/0 poisons the pixel with NaN in the reference), vanishing
    // cosines — takes the exact path below.  The power function was the longest dependent chain of the sorted
    // shading stage (tools/wave2_timing.py).
    {
        const float g = (float)ex * __log2f((float)rdotwi);
        if (rdotwi > 0.0 && fabsf(g) < 900.0f && fabs(ndotwi) > 1e-200 && fabs(ndotwi) < 1e200) {
            lobe = 1.0;
            weight = div_full(ndotwi, ndotwi);
            return;
        }
    }
#endif
    lobe = lobe_pow(rdotwi, ex);
    double pdf = lobe * ndotwi;
    weight = div_full(ndotwi, pdf);
}

__device__ __forceinline__ void glossy_sample(V3 normal, V3 dir, double sqx, double sqy, double ex, double inv_e1,
                                              V3 &wi, double &weight, double &lobe, bool &flipped) {
    glossy_sample_hs(normal, dir, to_unit_hemi_dev(sqx, sqy, inv_e1), ex, wi, weight, lobe, flipped);
}
