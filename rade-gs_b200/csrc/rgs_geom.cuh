// rgs_geom.cuh -- covariance geometry shared by forward and backward preprocess.
#pragma once
#include "rgs_common.cuh"

namespace rgs {

// ---- Jacobi eigen-solver for a symmetric 3x3 (general covariance path) ----------------------------
// Returns eigenvalues in lam[3] and eigenvectors as the columns of vec.  Cyclic sweeps; 6 sweeps reach
// float precision for any 3x3.
__device__ inline void eig_sym3_jacobi(const float cov[6], float lam[3], M3& vec) {
	float a00 = cov[0], a01 = cov[1], a02 = cov[2], a11 = cov[3], a12 = cov[4], a22 = cov[5];
	float v[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};  // v[row][col]
#pragma unroll 1
	for (int sweep = 0; sweep < 8; sweep++) {
		float off = fabsf(a01) + fabsf(a02) + fabsf(a12);
		float diag = fabsf(a00) + fabsf(a11) + fabsf(a22);
		if (off <= 1e-12f * diag || off == 0.f) break;
		// rotate (0,1)
		{
			if (a01 != 0.f) {
				float theta = (a11 - a00) / (2.f * a01);
				float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
				float c = 1.f / sqrtf(t * t + 1.f), s = t * c;
				float n00 = a00 - t * a01, n11 = a11 + t * a01;
				float n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
				a00 = n00; a11 = n11; a01 = 0.f; a02 = n02; a12 = n12;
#pragma unroll
				for (int r = 0; r < 3; r++) {
					float x = v[r][0], y = v[r][1];
					v[r][0] = c * x - s * y;
					v[r][1] = s * x + c * y;
				}
			}
		}
		// rotate (0,2)
		{
			if (a02 != 0.f) {
				float theta = (a22 - a00) / (2.f * a02);
				float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
				float c = 1.f / sqrtf(t * t + 1.f), s = t * c;
				float n00 = a00 - t * a02, n22 = a22 + t * a02;
				float n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
				a00 = n00; a22 = n22; a02 = 0.f; a01 = n01; a12 = n12;
#pragma unroll
				for (int r = 0; r < 3; r++) {
					float x = v[r][0], y = v[r][2];
					v[r][0] = c * x - s * y;
					v[r][2] = s * x + c * y;
				}
			}
		}
		// rotate (1,2)
		{
			if (a12 != 0.f) {
				float theta = (a22 - a11) / (2.f * a12);
				float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
				float c = 1.f / sqrtf(t * t + 1.f), s = t * c;
				float n11 = a11 - t * a12, n22 = a22 + t * a12;
				float n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
				a11 = n11; a22 = n22; a12 = 0.f; a01 = n01; a02 = n02;
#pragma unroll
				for (int r = 0; r < 3; r++) {
					float x = v[r][1], y = v[r][2];
					v[r][1] = c * x - s * y;
					v[r][2] = s * x + c * y;
				}
			}
		}
	}
	lam[0] = a00; lam[1] = a11; lam[2] = a22;
	vec = m3(v[0][0], v[1][0], v[2][0], v[0][1], v[1][1], v[2][1], v[0][2], v[1][2], v[2][2]);
}

// Inverse camera-space covariance applied to a vector, from an eigen-decomposition given in WORLD space:
// eigenvectors are the columns of E (world), eigenvalues lam.  Mirrors forward.cu:139-159:
//   well conditioned (lam_min > 1e-8):  Sigma^-1 = E diag(1/lam) E^T
//   otherwise:                           Sigma^-1 := e_min e_min^T
// returns Rv * Sigma^-1 * Rv^T * uvh  (cov_cam_inv * uvh, forward.cu:157-159).
__device__ __forceinline__ V3 apply_cov_cam_inv(const M3& E, const float lam[3], const float* V, V3 uvh, bool& well_conditioned, int& min_id) {
	min_id = lam[0] > lam[1] ? (lam[1] > lam[2] ? 2 : 1) : (lam[0] > lam[2] ? 2 : 0);
	well_conditioned = lam[min_id] > 0.00000001f;
	// a_k = Rv * e_k  (camera-space eigenvectors);  Rv[i][j] = V[i + 4 j]
	V3 a[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		V3 e = E.c[k];
		a[k] = V3{V[0] * e.x + V[4] * e.y + V[8] * e.z, V[1] * e.x + V[5] * e.y + V[9] * e.z, V[2] * e.x + V[6] * e.y + V[10] * e.z};
	}
	if (well_conditioned) {
		float w0 = dot3(a[0], uvh) / lam[0], w1 = dot3(a[1], uvh) / lam[1], w2 = dot3(a[2], uvh) / lam[2];
		return V3{a[0].x * w0 + a[1].x * w1 + a[2].x * w2, a[0].y * w0 + a[1].y * w1 + a[2].y * w2, a[0].z * w0 + a[1].z * w1 + a[2].z * w2};
	}
	V3 am = min_id == 0 ? a[0] : (min_id == 1 ? a[1] : a[2]);
	return am * dot3(am, uvh);
}

// Rotation matrix of quaternion (r,x,y,z), in the reference's column-major fill (forward.cu:286-290):
// the COLUMNS of the returned M3 are the ROWS of the usual rotation matrix.
__device__ __forceinline__ M3 quat_to_glm_rot(float r, float x, float y, float z) {
	return m3(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
	          2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
	          2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// Eigen-decomposition of Sigma for the geometry terms. Shared by forward and backward preprocess.
//   analytic (scale/rotation with unit quaternion): lam_k = (mod*s_k)^2, e_k = k-th column of the usual
//   rotation matrix = k-th ROW-collection of Rg, i.e. E = transpose(Rg).
__device__ __forceinline__ void sigma_eigen(bool analytic, const M3& Rg, V3 s_mod, const float cov3D[6], float lam[3], M3& E) {
	if (analytic) {
		lam[0] = s_mod.x * s_mod.x;
		lam[1] = s_mod.y * s_mod.y;
		lam[2] = s_mod.z * s_mod.z;
		E = transpose(Rg);
	} else {
		eig_sym3_jacobi(cov3D, lam, E);
	}
}

}  // namespace rgs
