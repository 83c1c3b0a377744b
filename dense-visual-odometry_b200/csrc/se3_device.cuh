// SE(3) bookkeeping and the 6x6 solve of one Gauss-Newton iteration, executed by ONE thread per pair.
//
// Mirrors the arithmetic of the reference's Lie classes as they behave in this image (NumPy 2 weak
// scalar promotion): quaternion wxyz + translation in float32, never renormalised; trigonometric
// scalars in float64.  Reference lines (relative to src/dense_visual_odometry/utils/lie_algebra/):
//   So3.__init__ (3,1) branch       special_orthogonal_group.py:33-50
//   So3._se3_to_quat                special_orthogonal_group.py:65-86
//   So3.exp / So3.log               special_orthogonal_group.py:158-209
//   quat_mult                       common.py:51-73
//   Se3.exp / log / inverse / mul   special_euclidean_group.py:35-96
//   Se3.from_se3                    special_euclidean_group.py:105-123
#pragma once
#include <math.h>

namespace dvo {

constexpr float kLieEps = 1e-6f;
constexpr float kPiF = 3.14159274101257324f;      // float32(np.pi)
constexpr float kTwoPiF = 6.28318548202514648f;   // float32(2*np.pi)

struct PoseQT {
    float q[4];  // w x y z
    float t[3];
};

__device__ __forceinline__ float norm3f(const float* v) {
    // np.linalg.norm on a float32 vector: sqrt of the float32 sum of squares
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
}

__device__ __forceinline__ float wrap_angle_f(float a) {
    // common.py:30-44 evaluated in float32: (a + pi) % (2 pi) - pi, Python modulo (sign of divisor)
    float s = __fadd_rn(a, kPiF);
    float m = fmodf(s, kTwoPiF);
    if (m < 0.0f) m = __fadd_rn(m, kTwoPiF);
    return __fadd_rn(m, -kPiF);
}

__device__ inline void quat_to_R(const float* q, float* R) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = __fadd_rn(__fmul_rn(2.0f, __fadd_rn(__fmul_rn(w, w), __fmul_rn(x, x))), -1.0f);
    R[1] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, y), -__fmul_rn(w, z)));
    R[2] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, z), __fmul_rn(w, y)));
    R[3] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, y), __fmul_rn(w, z)));
    R[4] = __fadd_rn(__fmul_rn(2.0f, __fadd_rn(__fmul_rn(w, w), __fmul_rn(y, y))), -1.0f);
    R[5] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(y, z), -__fmul_rn(w, x)));
    R[6] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(x, z), -__fmul_rn(w, y)));
    R[7] = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(y, z), __fmul_rn(w, x)));
    R[8] = __fadd_rn(__fmul_rn(2.0f, __fadd_rn(__fmul_rn(w, w), __fmul_rn(z, z))), -1.0f);
}

// So3.log: rotation vector of a (possibly unnormalised) quaternion.
__device__ inline void quat_log(const float* q, float* phi) {
    const float n = norm3f(q + 1);
    if (n < kLieEps) {
        phi[0] = phi[1] = phi[2] = 0.0f;
        return;
    }
    const float two_atan = (float)(2.0 * atan2((double)n, (double)q[0]));
    const float theta = wrap_angle_f(two_atan / n);
    phi[0] = theta * q[1];
    phi[1] = theta * q[2];
    phi[2] = theta * q[3];
}

// Se3.exp as a 3x4 row-major matrix [R|t]; R = I when the rotation vector is shorter than 1e-6.
__device__ inline void pose_matrix(const PoseQT& p, float* T) {
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    // Se3.exp uses R = I iff |So3.log| < 1e-6.  |log| = |wrap(2 atan2(n, w))| with n = |q_v|: for 0.5 < w < 1.5 and
    // n >= 1e-6 that is at least 2 atan(n / 1.5) > 1.3e-6, and for n < 1e-6 the log is defined as 0 -- so the
    // float64 atan2 is only needed for quaternions far from unit length or in the hemisphere w < 0.
    const float n = norm3f(p.q + 1);
    bool rotate;
    if (n < kLieEps) {
        rotate = false;
    } else if (p.q[0] > 0.5f && p.q[0] < 1.5f) {
        rotate = true;
    } else {
        float phi[3];
        quat_log(p.q, phi);
        rotate = norm3f(phi) >= kLieEps;
    }
    if (rotate) quat_to_R(p.q, R);
    T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = p.t[0];
    T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = p.t[1];
    T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = p.t[2];
}

// So3((3,1)): quaternion of a rotation vector plus the wrapped vector itself.
__device__ inline void phi_to_quat(const float* phi_in, float* q, float* phi_w) {
    const float theta = norm3f(phi_in);
    if (theta < kLieEps) {
        q[0] = 1.0f; q[1] = q[2] = q[3] = 0.0f;
        phi_w[0] = phi_w[1] = phi_w[2] = 0.0f;
        return;
    }
    const float a0 = phi_in[0] / theta, a1 = phi_in[1] / theta, a2 = phi_in[2] / theta;
    const float thw = wrap_angle_f(theta);
    phi_w[0] = thw * a0; phi_w[1] = thw * a1; phi_w[2] = thw * a2;
    const float th = norm3f(phi_w);
    const float x = phi_w[0] / th, y = phi_w[1] / th, z = phi_w[2] / th;
    double s, c;
    sincos((double)th * 0.5, &s, &c);
    q[0] = (float)c;
    q[1] = (float)(s * (double)x);
    q[2] = (float)(s * (double)y);
    q[3] = (float)(s * (double)z);
}

// Se3.from_se3: twist [v; w] (float32) -> pose.
__device__ inline void pose_from_xi(const float* xi, PoseQT& out) {
    float phi[3];
    phi_to_quat(xi + 3, out.q, phi);
    const float theta = norm3f(phi);
    if (theta < kLieEps) {
        out.q[0] = 1.0f; out.q[1] = out.q[2] = out.q[3] = 0.0f;
        out.t[0] = xi[0]; out.t[1] = xi[1]; out.t[2] = xi[2];
        return;
    }
    const double th = (double)theta;
    const float c1 = (float)((1.0 - cos(th)) / (th * th));
    const float c2 = (float)((th - sin(th)) / (th * th * th));
    // hat(phi) and its square, float32
    const float h[9] = {0.0f, -phi[2], phi[1], phi[2], 0.0f, -phi[0], -phi[1], phi[0], 0.0f};
    float V[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float h2 = 0.0f;
#pragma unroll
            for (int k = 0; k < 3; ++k) h2 = fmaf(h[i * 3 + k], h[k * 3 + j], h2);
            const float eye = (i == j) ? 1.0f : 0.0f;
            V[i * 3 + j] = __fadd_rn(__fadd_rn(eye, __fmul_rn(c1, h[i * 3 + j])), __fmul_rn(c2, h2));
        }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out.t[i] = fmaf(V[i * 3 + 2], xi[2], fmaf(V[i * 3 + 1], xi[1], V[i * 3 + 0] * xi[0]));
}

// Se3.__mul__: out = a * b (apply b first).
__device__ inline void pose_compose(const PoseQT& a, const PoseQT& b, PoseQT& out) {
    const float* qa = a.q;
    const float* qb = b.q;
    float q[4];
    q[0] = qa[0] * qb[0] - (qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3]);
    q[1] = qa[0] * qb[1] + qb[0] * qa[1] + (qa[2] * qb[3] - qa[3] * qb[2]);
    q[2] = qa[0] * qb[2] + qb[0] * qa[2] + (qa[3] * qb[1] - qa[1] * qb[3]);
    q[3] = qa[0] * qb[3] + qb[0] * qa[3] + (qa[1] * qb[2] - qa[2] * qb[1]);
    float R[9];
    quat_to_R(a.q, R);
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        t[i] = a.t[i] + fmaf(R[i * 3 + 2], b.t[2], fmaf(R[i * 3 + 1], b.t[1], R[i * 3 + 0] * b.t[0]));
#pragma unroll
    for (int i = 0; i < 4; ++i) out.q[i] = q[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) out.t[i] = t[i];
}

// So3._SE3_to_quat on R^T, then Se3.inverse.  The reference keeps this quaternion in float64 until the
// next product; float32 storage loses nothing the later float32 product would have kept.
__device__ inline void pose_inverse(const PoseQT& a, PoseQT& out) {
    float Rf[9];
    quat_to_R(a.q, Rf);
    double R[9];  // transpose
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = (double)Rf[j * 3 + i];
    double q[4] = {0, 0, 0, 0};
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double t = sqrt(1.0 + tr);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (R[7] - R[5]) * t;
        q[2] = (R[2] - R[6]) * t;
        q[3] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[1 + j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[1 + k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out.q[i] = (float)q[i];
    float Ri[9];
    quat_to_R(out.q, Ri);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out.t[i] = -fmaf(Ri[i * 3 + 2], a.t[2], fmaf(Ri[i * 3 + 1], a.t[1], Ri[i * 3 + 0] * a.t[0]));
}

// Se3.log: twist of a pose, float32 (used by the sigma prior only).
__device__ inline void pose_log(const PoseQT& p, float* xi) {
    float phi[3];
    quat_log(p.q, phi);
    const float theta = norm3f(phi);
    if (theta < kLieEps) {
        xi[0] = p.t[0]; xi[1] = p.t[1]; xi[2] = p.t[2];
        xi[3] = xi[4] = xi[5] = 0.0f;
        return;
    }
    float a[3] = {phi[0] / theta, phi[1] / theta, phi[2] / theta};
    float qa[4], aw[3];
    phi_to_quat(a, qa, aw);  // So3(a): unit axis goes through the wrap like any rotation vector
    const double th2 = (double)theta * 0.5;
    const double A = th2 * cos(th2) / sin(th2);
    const double ah[9] = {0.0, -aw[2], aw[1], aw[2], 0.0, -aw[0], -aw[1], aw[0], 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double vinv = ((i == j) ? A : 0.0) + (1.0 - A) * (double)aw[i] * (double)aw[j] - th2 * ah[i * 3 + j];
            acc += vinv * (double)p.t[j];
        }
        xi[i] = (float)acc;
    }
    xi[3] = phi[0]; xi[4] = phi[1]; xi[5] = phi[2];
}

// Solves H x = b for the symmetric 6x6 H (float64, LDL^T without pivoting).  A pivot that is not
// safely positive relative to the largest diagonal entry drops that unknown (x_i = 0), which is what
// the reference's minimum-norm gelsy solve (base_robust_dvo.py:196-198, rcond = eps(float32)) does
// for an exactly decoupled degenerate direction.  Returns the number of dropped unknowns.
// Fully unrolled (everything stays in registers): a dropped pivot gets d = 1/d = 0, which zeroes its
// column of L and its unknown without any further branching.
__device__ __forceinline__ int solve6_ldlt(const double* Hin, const double* b, double* x) {
    double L[6][6];
    double d[6], dinv[6];
    double dmax = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) dmax = fmax(dmax, fabs(Hin[i * 6 + i]));
    const double tiny = dmax * 1.1920929e-07;
    int ndrop = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double dj = Hin[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k] * d[k];
        const bool ok = dj > tiny;
        ndrop += ok ? 0 : 1;
        d[j] = ok ? dj : 0.0;
        dinv[j] = ok ? 1.0 / dj : 0.0;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double sum = Hin[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) sum -= L[i][k] * L[j][k] * d[k];
            L[i][j] = sum * dinv[j];
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sum = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sum -= L[i][k] * y[k];
        y[i] = (dinv[i] != 0.0) ? sum : 0.0;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sum = y[i] * dinv[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) sum -= L[k][i] * x[k];
        x[i] = sum;
    }
    return ndrop;
}

}  // namespace dvo
