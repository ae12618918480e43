// rgs_geom.cuh -- covariance geometry shared by forward and backward preprocess.
//
// The ray-space plane / normal of a splat comes from Sigma^-1, which the reference builds from a 3x3 symmetric
// eigen-decomposition (cuda_rasterizer/forward.cu:135-159, backward.cu:221-247) computed by a Householder
// tridiagonalisation followed by implicit-shift QL sweeps whose convergence tests are ABSOLUTE (|x| <= 1e-7,
// cuda_rasterizer/auxiliary.h:182-401 -- glm's findEigenvaluesSymReal with its relative tests replaced).
// For splats with sigma ~ 1e-2 (covariance entries ~ 1e-4) that absolute test stops the sweeps with
// eigenvectors only good to ~1e-3 rad, and those eigenvectors shape the rendered normals and every geometry
// gradient.  Parity with the reference therefore needs the same algorithm with the same stopping rules, not a
// more accurate one: eig_sym3_tql below follows the published tred2/tqli scheme specialised to 3x3 with those
// absolute tests, operation by operation (divisions and square roots spelled as the IEEE intrinsics so the
// result is independent of per-file fast-division settings).  (An exact decomposition R S^2 R^T is available for free when
// Sigma comes from scale+rotation; measured on C2 it moves 5.6 % of the normal map by more than 1e-4, so it
// is not used.)
#pragma once
#include "rgs_common.cuh"

namespace rgs {

constexpr float kEigEps = 0.0000001f;

__device__ __forceinline__ float hypot_nr(float a, float b) {  // sqrt(a^2+b^2) without overflow (auxiliary.h:200-214)
	float absa = fabsf(a), absb = fabsf(b);
	if (absa > absb) {
		absb = __fdiv_rn(absb, absa);
		absb *= absb;
		return absa * __fsqrt_rn(1.f + absb);
	}
	if (fabsf(absb) <= kEigEps) return 0.f;
	absa = __fdiv_rn(absa, absb);
	absa *= absa;
	return absb * __fsqrt_rn(1.f + absa);
}

// One implicit-shift QL sweep on the unreduced block [L..M] of the tridiagonal (d, e), rotations accumulated in z
// (auxiliary.h:355-388).  L and M are compile-time so every array index is static.
template <int L, int M>
__device__ __forceinline__ void ql_sweep(float (&d)[3], float (&e)[3], float (&z)[3][3]) {
	float g = __fdiv_rn(d[L + 1] - d[L], 2 * e[L]);
	float r = hypot_nr(g, 1.f);
	g = d[M] - d[L] + __fdiv_rn(e[L], g + ((g >= 0.f) ? fabsf(r) : -fabsf(r)));
	float s = 1.f, c = 1.f, p = 0.f;
	bool underflow = false;  // the reference's early `break` (r ~ 0)
#pragma unroll
	for (int i = M - 1; i >= L; i--) {
		if (!underflow) {
			float f = s * e[i];
			const float b = c * e[i];
			e[i + 1] = r = hypot_nr(f, g);
			if (fabsf(r) <= kEigEps) {
				d[i + 1] -= p;
				e[M] = 0.f;
				underflow = true;
			} else {
				s = __fdiv_rn(f, r);
				c = __fdiv_rn(g, r);
				g = d[i + 1] - p;
				r = (d[i] - g) * s + 2 * c * b;
				d[i + 1] = g + (p = s * r);
				g = c * r - b;
#pragma unroll
				for (int k = 0; k < 3; k++) {
					f = z[k][i + 1];
					z[k][i + 1] = s * z[k][i] + c * f;
					z[k][i] = c * z[k][i] - s * f;
				}
			}
		}
	}
	if (!underflow) {
		d[L] -= p;
		e[L] = g;
		e[M] = 0.f;
	}
}

// Eigen-decomposition of the symmetric matrix with packed upper triangle cov = (xx,xy,xz,yy,yz,zz).
// lam[k] / column k of `vec` are an eigenpair.  Returns false if a QL sweep did not converge in 30 iterations
// (the reference then zeroes the geometry terms, forward.cu:162-168).
__device__ inline bool eig_sym3_tql(const float cov[6], float lam[3], M3& vec) {
	// z[r][c]: work matrix, ends up holding the eigenvectors in its columns
	float z[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
	float d[3], e[3];

	// ---- reduction to tridiagonal form: for 3x3 a single Householder reflection built from row 2 ----
	{
		float h = 0.f;
		const float scale = fabsf(z[2][0]) + fabsf(z[2][1]);
		if (fabsf(scale) <= kEigEps) {
			e[2] = z[2][1];
		} else {
			z[2][0] = __fdiv_rn(z[2][0], scale);
			h += z[2][0] * z[2][0];
			z[2][1] = __fdiv_rn(z[2][1], scale);
			h += z[2][1] * z[2][1];
			float f = z[2][1];
			float g = (f >= 0.f) ? -__fsqrt_rn(h) : __fsqrt_rn(h);
			e[2] = scale * g;
			h -= f * g;
			z[2][1] = f - g;
			f = 0.f;
			// j = 0
			z[0][2] = __fdiv_rn(z[2][0], h);
			g = 0.f;
			g += z[0][0] * z[2][0];
			g += z[1][0] * z[2][1];
			e[0] = __fdiv_rn(g, h);
			f += e[0] * z[2][0];
			// j = 1
			z[1][2] = __fdiv_rn(z[2][1], h);
			g = 0.f;
			g += z[1][0] * z[2][0];
			g += z[1][1] * z[2][1];
			e[1] = __fdiv_rn(g, h);
			f += e[1] * z[2][1];
			const float hh = __fdiv_rn(f, h + h);
			// j = 0
			f = z[2][0];
			e[0] = g = e[0] - hh * f;
			z[0][0] -= (f * e[0] + g * z[2][0]);
			// j = 1
			f = z[2][1];
			e[1] = g = e[1] - hh * f;
			z[1][0] -= (f * e[0] + g * z[2][0]);
			z[1][1] -= (f * e[1] + g * z[2][1]);
		}
		d[2] = h;
		e[1] = z[1][0];
		d[1] = 0.f;
		d[0] = 0.f;
		e[0] = 0.f;
	}
	// ---- accumulate the transformation ----
	d[0] = z[0][0];
	z[0][0] = 1.f;
	d[1] = z[1][1];
	z[1][1] = 1.f;
	z[0][1] = z[1][0] = 0.f;
	if (!(fabsf(d[2]) <= kEigEps)) {
#pragma unroll
		for (int j = 0; j < 2; j++) {
			float g = 0.f;
			g += z[2][0] * z[0][j];
			g += z[2][1] * z[1][j];
			z[0][j] -= g * z[0][2];
			z[1][j] -= g * z[1][2];
		}
	}
	d[2] = z[2][2];
	z[2][2] = 1.f;
	z[0][2] = z[2][0] = 0.f;
	z[1][2] = z[2][1] = 0.f;

	// ---- implicit QL on the tridiagonal (d, e) ----
	// The reference loops l = 0..2 with run-time m and i; here l is unrolled and m dispatched through templates so
	// that d, e, z are indexed statically and stay in registers (same operations in the same order).
	e[0] = e[1];
	e[1] = e[2];
	e[2] = 0.f;
	{  // l = 0
		int iter = 0;
		while (true) {
			const int m = (fabsf(fabsf(e[0])) <= kEigEps) ? 0 : ((fabsf(fabsf(e[1])) <= kEigEps) ? 1 : 2);
			if (m == 0) break;
			if (iter++ == 30) return false;
			if (m == 1) ql_sweep<0, 1>(d, e, z); else ql_sweep<0, 2>(d, e, z);
		}
	}
	{  // l = 1
		int iter = 0;
		while (true) {
			if (fabsf(fabsf(e[1])) <= kEigEps) break;
			if (iter++ == 30) return false;
			ql_sweep<1, 2>(d, e, z);
		}
	}
	lam[0] = d[0];
	lam[1] = d[1];
	lam[2] = d[2];
	vec = m3(z[0][0], z[1][0], z[2][0], z[0][1], z[1][1], z[2][1], z[0][2], z[1][2], z[2][2]);
	return true;
}

// Sigma^-1 substitute of the reference (forward.cu:139-155): E diag(1/lam) E^T when the smallest eigenvalue is
// above 1e-8, else the rank-1 projector on the smallest eigenvector.
struct SigmaInv {
	M3 inv;          // Vrk_inv
	M3 E;            // eigenvectors (columns)
	float lam[3];
	float lam_min;
	int min_id;
	bool well;
	bool solved;     // eigen-solver converged
};

__device__ __forceinline__ SigmaInv sigma_inverse(const float cov3D[6]) {
	SigmaInv s;
	s.solved = eig_sym3_tql(cov3D, s.lam, s.E);
	const float* l = s.lam;
	s.min_id = l[0] > l[1] ? (l[1] > l[2] ? 2 : 1) : (l[0] > l[2] ? 2 : 0);
	s.lam_min = s.min_id == 0 ? l[0] : (s.min_id == 1 ? l[1] : l[2]);
	s.well = s.lam_min > 0.00000001f;
	if (s.well) {
		const M3 diag = m3(__fdiv_rn(1.f, l[0]), 0.f, 0.f, 0.f, __fdiv_rn(1.f, l[1]), 0.f, 0.f, 0.f, __fdiv_rn(1.f, l[2]));
		s.inv = s.E * diag * transpose(s.E);
	} else {
		const V3 em = s.min_id == 0 ? s.E.c[0] : (s.min_id == 1 ? s.E.c[1] : s.E.c[2]);
		s.inv = outer(em, em);
	}
	return s;
}

// Rotation matrix of quaternion (r,x,y,z), in the reference's column-major fill (forward.cu:286-290):
// the COLUMNS of the returned M3 are the ROWS of the usual rotation matrix.
__device__ __forceinline__ M3 quat_to_glm_rot(float r, float x, float y, float z) {
	return m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
	          2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
	          2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

}  // namespace rgs
