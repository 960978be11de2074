// curve.cuh — BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Device-side replacement for halo2curves::bn256::{G1Affine, G1} (reference import:
// halo2-circuits/src/ecc/ecdsa_p256.rs:27).  Affine points are halo2curves' in-memory G1Affine
// ({x,y} Montgomery, identity = (0,0)).  Accumulators use extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity ZZ = 0): a mixed addition is 8M + 2S with no
// inversion, which is what the bucket accumulation is made of.  Formulas: EFD short-Weierstrass
// xyzz, a = 0 (madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1).
#pragma once
#include "field.cuh"

namespace zkw {

struct G1Affine {
    Fq x, y;
    __host__ __device__ __forceinline__ bool is_identity() const { return x.is_zero() && y.is_zero(); }
    __host__ __device__ __forceinline__ static G1Affine load(const void* p) {
        G1Affine r;
        r.x = Fq::load(p);
        r.y = Fq::load(reinterpret_cast<const char*>(p) + 32);
        return r;
    }
    __device__ __forceinline__ static G1Affine load_nc(const void* p) {
        G1Affine r;
        r.x = Fq::load_nc(p);
        r.y = Fq::load_nc(reinterpret_cast<const char*>(p) + 32);
        return r;
    }
    __host__ __device__ __forceinline__ void store(void* p) const {
        x.store(p);
        y.store(reinterpret_cast<char*>(p) + 32);
    }
};

struct G1Xyzz {
    Fq x, y, zz, zzz;

    __host__ __device__ __forceinline__ static G1Xyzz identity() {
        G1Xyzz r;
        r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
        return r;
    }
    __host__ __device__ __forceinline__ bool is_identity() const { return zz.is_zero(); }
    __host__ __device__ __forceinline__ static G1Xyzz from_affine(const G1Affine& a) {
        G1Xyzz r;
        if (a.is_identity()) return identity();
        r.x = a.x; r.y = a.y; r.zz = Fq::one(); r.zzz = Fq::one();
        return r;
    }
    __host__ __device__ __forceinline__ static G1Xyzz load(const void* p) {
        const char* c = reinterpret_cast<const char*>(p);
        G1Xyzz r;
        r.x = Fq::load(c); r.y = Fq::load(c + 32); r.zz = Fq::load(c + 64); r.zzz = Fq::load(c + 96);
        return r;
    }
    __device__ __forceinline__ static G1Xyzz load_cg(const void* p) {
        const char* c = reinterpret_cast<const char*>(p);
        G1Xyzz r;
        r.x = Fq::load_cg(c); r.y = Fq::load_cg(c + 32); r.zz = Fq::load_cg(c + 64); r.zzz = Fq::load_cg(c + 96);
        return r;
    }
    __host__ __device__ __forceinline__ void store(void* p) const {
        char* c = reinterpret_cast<char*>(p);
        x.store(c); y.store(c + 32); zz.store(c + 64); zzz.store(c + 96);
    }

    // 2 * (affine point), a != identity
    __host__ __device__ __forceinline__ static G1Xyzz double_affine(const G1Affine& a) {
        G1Xyzz r;
        Fq u = a.y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = a.x * v;
        Fq xx = a.x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * a.y;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    __host__ __device__ __forceinline__ G1Xyzz dbl() const {
        if (is_identity()) return *this;
        G1Xyzz r;
        Fq u = y.dbl();
        Fq v = u.sqr();
        Fq w = u * v;
        Fq s = x * v;
        Fq xx = x.sqr();
        Fq m = xx.dbl() + xx;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += b (affine); `negate` adds -b instead.
    __host__ __device__ __forceinline__ void add_mixed(const G1Affine& b_in, bool negate = false) {
        if (b_in.is_identity()) return;
        G1Affine b = b_in;
        if (negate) b.y = b.y.neg();
        if (is_identity()) { *this = from_affine(b); return; }
        Fq u2 = b.x * zz;
        Fq s2 = b.y * zzz;
        Fq p = u2 - x;
        Fq r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = double_affine(b);
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = x * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - y * ppp;
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += b (affine, canonical coordinates) with the running point's coordinates lazily reduced (in [0, 2p),
    // see field.cuh): same formulas as add_mixed, ten products without their final conditional subtraction.
    // The result is again in [0, 2p); call normalize() before the point leaves the kernel.
    __device__ __forceinline__ void add_mixed_lazy(const G1Affine& b_in, bool negate) {
        if (b_in.is_identity()) return;
        G1Affine b = b_in;
        if (negate) b.y = b.y.neg();
        if (is_identity()) { *this = from_affine(b); return; }
        Fq u2 = Fq::mul_lazy(b.x, zz);
        Fq s2 = Fq::mul_lazy(b.y, zzz);
        Fq p = Fq::sub_lazy(u2, x);
        Fq r = Fq::sub_lazy(s2, y);
        if (p.is_zero_lazy()) {   // same x: doubling or cancellation, through the canonical formulas
            normalize();
            add_mixed(b, false);
            return;
        }
#ifdef ZKW_MSM_NO_SQR
        Fq pp = Fq::mul_lazy(p, p);
        const Fq rr = Fq::mul_lazy(r, r);
#else
        Fq pp = Fq::sqr_lazy(p);          // the two squarings of the 8M + 2S addition: 100 wide multiplies instead of 128 each
        const Fq rr = Fq::sqr_lazy(r);
#endif
        Fq ppp = Fq::mul_lazy(p, pp);
        Fq q = Fq::mul_lazy(x, pp);
        Fq x3 = Fq::sub_lazy(Fq::sub_lazy(rr, ppp), Fq::add_lazy(q, q));
#ifdef ZKW_MSM_NO_MUL2
        y = Fq::sub_lazy(Fq::mul_lazy(r, Fq::sub_lazy(q, x3)), Fq::mul_lazy(y, ppp));
#else
        // Y3 = R (Q - X3) + (-Y1) PPP as one two-product Montgomery pass (one reduction instead of two: 192 wide multiplies for
        // 256), result < 2.52 p brought back below 2p
        y = Fq::mul2_lazy(r, Fq::sub_lazy(q, x3), y.neg_2m(), ppp).reduced_2m();
#endif
        x = x3;
        zz = Fq::mul_lazy(zz, pp);
        zzz = Fq::mul_lazy(zzz, ppp);
    }
    __host__ __device__ __forceinline__ void normalize() {
        x = x.normalized(); y = y.normalized(); zz = zz.normalized(); zzz = zzz.normalized();
    }

    // this += b (XYZZ)
    __host__ __device__ __forceinline__ void add(const G1Xyzz& b) {
        if (b.is_identity()) return;
        if (is_identity()) { *this = b; return; }
        Fq u1 = x * b.zz;
        Fq u2 = b.x * zz;
        Fq s1 = y * b.zzz;
        Fq s2 = b.y * zzz;
        Fq p = u2 - u1;
        Fq r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        Fq pp = p.sqr();
        Fq ppp = p * pp;
        Fq q = u1 * pp;
        Fq x3 = r.sqr() - ppp - q.dbl();
        y = r * (q - x3) - s1 * ppp;
        x = x3;
        zz = zz * b.zz * pp;
        zzz = zzz * b.zzz * ppp;
    }

    // Jacobian (X', Y', Z') with Z' = ZZZ: X' = X*ZZ^2, Y' = Y*ZZZ^2 (x = X'/Z'^2, y = Y'/Z'^3)
    __host__ __device__ __forceinline__ void to_jacobian(Fq& jx, Fq& jy, Fq& jz) const {
        if (is_identity()) { jx = Fq::zero(); jy = Fq::one(); jz = Fq::zero(); return; }
        jx = x * zz.sqr();
        jy = y * zzz.sqr();
        jz = zzz;
    }
};

}  // namespace zkw
