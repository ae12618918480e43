// rgs_render_bwd.cu -- per-tile backward of the alpha blend for sm_100a.
//
// Replaces renderCUDA<3,COORD,DEPTH,NORMAL> backward (reference: cuda_rasterizer/backward.cu:631-1016,
// dispatch :1101-1163).  The per-pixel recurrence (back-to-front re-traversal, T recovered by division,
// running suffix blends, median terms) is the reference's; the gradient scatter is not:
//   * the reference issues 10/16/22/25 scalar global atomicAdds per contributing (pixel, splat) pair
//     (backward.cu:878-1013).  Here the 32 pixels of a warp evaluate the same splat in lock-step, the
//     16 (or 32) per-splat partial gradients are combined with a recursive-halving shuffle
//     reduce-scatter (16 resp. 31 SHFL instead of 80 resp. 125 for a butterfly per value), and ONE
//     coalesced red.global.add of 64 B (128 B) per (warp, splat) lands in a packed accumulator row;
//   * constant factors (0.5*W, 0.5*H, 1/focal) are pulled out of the sums and applied once per Gaussian
//     in backward-preprocess;
//   * the tile list is cut at the block-wide maximum of last_contributor (nothing behind it can receive
//     gradient, backward.cu:838-839), each warp additionally skips splats behind ITS maximum and splats
//     that cannot reach alpha >= 1/255 inside its 8x4 pixel block -- read from the ballots forward-render stored;
//   * records are gathered with cp.async into a 2-stage ring, like forward, but walking the list from
//     the back.
#include "rgs_render_common.cuh"

namespace rgs {

namespace {

// Recursive-halving reduce-scatter over a warp.  On return lane L holds, in v[0], the warp-wide sum of
// value index (L >> 1) for N == 16 (both lanes of a pair hold it) or index L for N == 32.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane) {
	static_assert(N == 16 || N == 32, "16 or 32 values");
#pragma unroll
	for (int h = N / 2, mask = 16; h >= 2; h >>= 1, mask >>= 1) {
		const bool upper = (lane & mask) != 0;
#pragma unroll
		for (int k = 0; k < h; k += 2) {
			// two values per step: the adds go through the packed fp32x2 pipe (one FADD2 instead of two FADD)
			const float s0 = upper ? v[k] : v[k + h], s1 = upper ? v[k + 1] : v[k + h + 1];
			const float2 keep = make_float2(upper ? v[k + h] : v[k], upper ? v[k + h + 1] : v[k + 1]);
			const float2 recv = make_float2(__shfl_xor_sync(0xffffffffu, s0, mask), __shfl_xor_sync(0xffffffffu, s1, mask));
			const float2 sum = __fadd2_rn(keep, recv);
			v[k] = sum.x;
			v[k + 1] = sum.y;
		}
	}
	{
		const int mask = (N == 16) ? 2 : 1;
		const bool upper = (lane & mask) != 0;
		const float send = upper ? v[0] : v[1];
		const float keep = upper ? v[1] : v[0];
		v[0] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
	}
	if (N == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float rcp_approx(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
	int v;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}

}  // namespace

template <bool COORD, bool DEPTH>
__global__ void __launch_bounds__(NTHREADS, (COORD ? 3 : 4)) render_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, const float* __restrict__ records,
    int W, int H, int py_off, int Hs, int grid_x, int row_begin, float focal_x, float focal_y, const float* __restrict__ bg_color,
    const float* __restrict__ alphas, const float* __restrict__ normalmap, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ accum_depth, const float* __restrict__ accum_coord, const float* __restrict__ normal_length,
    const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_coords, const float* __restrict__ dL_dpixel_mcoords,
    const float* __restrict__ dL_dpixel_depths, const float* __restrict__ dL_dpixel_mdepths, const float* __restrict__ dL_dalphas,
    const float* __restrict__ dL_dpixel_normals, float* __restrict__ grad_accum, const uint32_t* __restrict__ chunk_base,
    const uint32_t* __restrict__ hitmask) {
	constexpr bool GEO = COORD || DEPTH;
	constexpr int RFQ = COORD ? 6 : 4;
	constexpr int GF = COORD ? GRAD_FLOATS_COORD : GRAD_FLOATS_BASE;
	extern __shared__ float4 smem[];  // [2][RFQ][BATCH] records, then [2][BATCH] ids
	int* s_ids = reinterpret_cast<int*>(smem + 2 * RFQ * BATCH);
	const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
	__shared__ int s_block_last[NTHREADS / 32];

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int tile_x = blockIdx.x, tile_y = blockIdx.y + row_begin;
	const int bx0 = tile_x * TILE_X + (warp & 1) * 8, by0 = tile_y * TILE_Y + (warp >> 1) * 4;
	const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float wx0 = (float)bx0, wx1 = (float)min(bx0 + 7, W - 1), wy0 = (float)by0, wy1 = (float)min(by0 + 3, H - 1);
	const int pix_id = W * (py - py_off) + px;  // maps hold pixel rows [py_off, py_off + Hs)
	const size_t HW = (size_t)Hs * W;

	const uint2 range = ranges[tile_y * grid_x + tile_x];
	const uint32_t* my_mask = hitmask + (size_t)chunk_base[tile_y * grid_x + tile_x] * 8 + warp;

	// ---- per-pixel state from forward (backward.cu:704-781) ----
	const float T_final = inside ? (1 - alphas[pix_id]) : 0;
	const float w_final = inside ? alphas[pix_id] : 0;
	float T = T_final;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
	const int max_contributor = inside ? (int)n_contrib[pix_id + HW] : 0;

	float dL_dpixel[3] = {0.f, 0.f, 0.f};
	float dL_dpixel_coord[3] = {0.f, 0.f, 0.f}, dL_dpixel_mcoord[3] = {0.f, 0.f, 0.f};
	float dL_dpixel_t = 0.f, dL_dpixel_mt = 0.f, dL_dalpha = 0.f;
	float dL_dpixel_normal[3] = {0.f, 0.f, 0.f};
	if (inside) {
#pragma unroll
		for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * HW + pix_id];
		dL_dalpha = dL_dalphas[pix_id];
		if constexpr (GEO) {
			const float ww = w_final * w_final;
			if constexpr (COORD) {
#pragma unroll
				for (int i = 0; i < 3; i++) {
					const float g = dL_dpixel_coords[i * HW + pix_id];
					dL_dalpha -= g * accum_coord[i * HW + pix_id] / ww;
					dL_dpixel_coord[i] = g / w_final;
					dL_dpixel_mcoord[i] = dL_dpixel_mcoords[i * HW + pix_id];
				}
			}
			if constexpr (DEPTH) {
				const float pnx = (pxf - W / 2.f) / focal_x, pny = (pyf - H / 2.f) / focal_y;
				const float ln = sqrt(pnx * pnx + pny * pny + 1);
				const float g = dL_dpixel_depths[pix_id];
				dL_dalpha -= g * accum_depth[pix_id] / ww;
				dL_dpixel_t = g / w_final / ln;
				dL_dpixel_mt = dL_dpixel_mdepths[pix_id] / ln;
			}
			{
				const float gx = dL_dpixel_normals[pix_id], gy = dL_dpixel_normals[HW + pix_id], gz = dL_dpixel_normals[2 * HW + pix_id];
				const float nx = normalmap[pix_id], ny = normalmap[HW + pix_id], nz = normalmap[2 * HW + pix_id];
				const float nlen = normal_length[pix_id];
				if (nlen < 1.0E-12F) {
					dL_dpixel_normal[0] = gx / 1.0E-12F;
					dL_dpixel_normal[1] = gy / 1.0E-12F;
					dL_dpixel_normal[2] = gz / 1.0E-12F;
				} else {
					const float d = gx * nx + gy * ny + gz * nz;
					dL_dpixel_normal[0] = (gx - d * nx) / nlen;
					dL_dpixel_normal[1] = (gy - d * ny) / nlen;
					dL_dpixel_normal[2] = (gz - d * nz) / nlen;
				}
			}
		}
	}
	if (last_contributor == 0) {
		// pixels nothing was blended into: w_final = 0 makes the normalisation terms 0/0.  They never receive a splat
		// (`ok` is false for every pair) but the straight-line body multiplies by them, so clear them.
		dL_dalpha = 0.f; dL_dpixel_t = 0.f; dL_dpixel_mt = 0.f;
#pragma unroll
		for (int i = 0; i < 3; i++) { dL_dpixel[i] = 0.f; dL_dpixel_coord[i] = 0.f; dL_dpixel_mcoord[i] = 0.f; dL_dpixel_normal[i] = 0.f; }
	}
	float bg_dot_dpixel = 0;
#pragma unroll
	for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_dpixel[i];

	// ---- cut the list at the last splat that contributed to any pixel of the block / warp ----
	int warp_last = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
	if (lane == 0) s_block_last[warp] = warp_last;
	__syncthreads();
	int n = 0;
#pragma unroll
	for (int w = 0; w < NTHREADS / 32; w++) n = max(n, s_block_last[w]);
	n = min(n, (int)(range.y - range.x));  // defensive; last_contributor <= list length by construction
	const int rounds = (n + BATCH - 1) / BATCH;

	// running sum over the splats behind the current one (see the loop body); starts with the background layer
	float B = T_final * bg_dot_dpixel;
	// per-pixel upstream gradients in the pairing of the record fields (see the loop body)
	const float2 g_rg = make_float2(dL_dpixel[0], dL_dpixel[1]);
	const float2 g_bt = make_float2(dL_dpixel[2], dL_dpixel_t);
	const float2 g_nxy = make_float2(dL_dpixel_normal[0], dL_dpixel_normal[1]);
	const float g_nz = dL_dpixel_normal[2];
	const float2 npix = make_float2(-pxf, -pyf);

	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;

	// staged index j of round i is list position  n-1 - (i*BATCH + j)
	const size_t rec_stride = (size_t)RFQ * 4;
	auto issue_gather = [&](int stage, int id) {
		if (id >= 0) {
			const float4* src = reinterpret_cast<const float4*>(records + (size_t)id * rec_stride);
			float4* dst = smem + (size_t)stage * RFQ * BATCH + tid;
#pragma unroll
			for (int c = 0; c < RFQ; c++) cp_async16(dst + c * BATCH, src + c);
		}
		s_ids[stage * BATCH + tid] = id;
		cp_async_commit();
	};
	int id_cur = (tid < n) ? (int)point_list[range.x + n - 1 - tid] : -1;
	issue_gather(0, id_cur);
	int id_next = (BATCH + tid < n) ? (int)point_list[range.x + n - 1 - BATCH - tid] : -1;

	for (int i = 0; i < rounds; i++) {
		cp_async_wait_all();
		__syncthreads();
		if (i + 1 < rounds) {
			issue_gather((i + 1) & 1, id_next);
			const int nxt = (i + 2) * BATCH + tid;
			id_next = (nxt < n) ? (int)point_list[range.x + n - 1 - nxt] : -1;
		}
		const float4* s = smem + (size_t)(i & 1) * RFQ * BATCH;
		const uint32_t s_addr = smem_base + (uint32_t)((i & 1) * RFQ * BATCH * 16);
		const uint32_t ids_addr = smem_base + (uint32_t)(2 * RFQ * BATCH * 16 + (i & 1) * BATCH * 4);
		const int cnt = min(BATCH, n - i * BATCH);
		const int pos0 = n - 1 - i * BATCH;  // list position of staged index 0
		if (pos0 - (cnt - 1) >= warp_last) continue;  // whole batch is behind this warp's last contributor

		// hit masks of forward-render for the positions of this batch: word k covers list positions [32k, 32k+32);
		// lane l fetches word (pos0>>5) - l, chunks below pick theirs up with a shuffle
		const int k_hi = pos0 >> 5;
		uint32_t wl = 0;
		if (lane <= 9 && k_hi - lane >= 0) wl = __ldg(my_mask + (size_t)(k_hi - lane) * 8);

		for (int c0 = 0; c0 < cnt; c0 += 32) {
			// staged index j of this chunk <-> list position hi - j
			const int hi = pos0 - c0, lo_pos = hi - 31;
			const int k0 = lo_pos >> 5;  // floor, -1 when the chunk reaches below position 0
			const uint32_t w0 = __shfl_sync(0xffffffffu, wl, (k_hi - k0) & 31);
			const uint32_t w1 = __shfl_sync(0xffffffffu, wl, (k_hi - k0 - 1) & 31);
			unsigned m = __brev(__funnelshift_r(w0, w1, lo_pos & 31));  // bit j <-> position hi - j
			const int valid = min(32, cnt - c0);                       // staged entries in this chunk
			if (valid < 32) m &= (1u << valid) - 1u;
			const int behind = hi - warp_last + 1;                      // entries j < behind sit at or behind warp_last
			if (behind > 0) m = behind >= 32 ? 0u : (m & ~((1u << behind) - 1u));
			while (m) {
				const int bpos = __ffs(m) - 1;
				m &= m - 1;
				const int jj = c0 + bpos;
				const int contributor = pos0 - jj;  // 0-based position in the tile list (backward.cu:837)
				const uint32_t sa = s_addr + (uint32_t)jj * 16u;
				const float4 q0 = lds128(sa), q1 = lds128(sa + BATCH * 16u);  // (mx my A B) (C o ray.x ray.y)
				const float2 d = __fadd2_rn(make_float2(q0.x, q0.y), npix);   // centre - pixel
				const float dx = d.x, dy = d.y;
				const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
				const float G = expf(power);
				const float alpha_raw = min(0.99f, q1.y * G);
				const bool ok = contributor < last_contributor && !(power > 0.0f) && !(alpha_raw < 1.0f / 255.0f);
				if (!__any_sync(0xffffffffu, ok)) continue;

				// Straight-line body.  A lane whose pixel does not receive this splat runs it with alpha = 0: a layer of
				// zero opacity leaves T and the running sum below unchanged, its channel weights are alpha*T = 0, and the
				// terms that would not vanish (they all carry the factor G) are masked through G.
				//
				// Blend backward in "dot-product" form.  The reference keeps, per channel, the normalised blend of everything
				// behind the current splat (accum_rec, backward.cu:870-962) and forms sum_ch (c_ch - accum_rec_ch) g_ch * T.
				// With B = sum over the splats behind of (alpha T) * S, S = sum_ch c_ch g_ch (g = the pixel's upstream
				// gradients, alpha channel c = 1, background folded in as the last layer) that is exactly
				//      dL/dalpha = T * S - B / (1 - alpha),       B += alpha T S
				// -- one running scalar instead of 8 running blends + 8 "last" values (same algebra, different rounding).
				//
				// Channels that share a factor are kept in aligned register pairs -- (r,g) (b,t) (nx,ny) on the record side,
				// (g_r,g_g) (g_b,g_t) (g_nx,g_ny) on the pixel side -- and go through the packed fp32x2 FMUL2 / FFMA2.
				const float alpha = ok ? alpha_raw : 0.f;
				float4 q2 = lds128(sa + 2 * BATCH * 16u);  // r g b t_center
				float gv[GF];
				const float inv = rcp_approx(1.f - alpha);  // 1 - alpha in [0.01, 1]: bare MUFU.RCP; shared by T/(1-alpha) and B/(1-alpha)
				T = T * inv;
				const float w = alpha * T;  // dchannel_dcolor
				const float2 w2 = make_float2(w, w);
				if constexpr (DEPTH) q2.w = q2.w + (q1.z * dx + q1.w * dy);  // t = t_center + ray . d
				float2 S2 = __ffma2_rn(make_float2(q2.x, q2.y), g_rg, make_float2(dL_dalpha, 0.f));
				S2 = __ffma2_rn(make_float2(q2.z, q2.w), g_bt, S2);
				const float2 c_rg = __fmul2_rn(w2, g_rg);
				float2 c_bt = __fmul2_rn(w2, g_bt);  // (dL_dcolor.b, dL_dt before the median term)
				gv[G_COL + 0] = c_rg.x;
				gv[G_COL + 1] = c_rg.y;
				gv[G_COL + 2] = c_bt.x;
				float dL_dcoords[3] = {0.f, 0.f, 0.f};
				float4 q3, q4, q5;
				const bool is_median = ok && contributor == max_contributor - 1;
				if constexpr (GEO) q3 = lds128(sa + 3 * BATCH * 16u);
				if constexpr (COORD) {
					q4 = lds128(sa + 4 * BATCH * 16u);
					q5 = lds128(sa + 5 * BATCH * 16u);
					const float coord[3] = {q4.x + q4.w * dx + q5.x * dy, q4.y + q5.y * dx + q5.z * dy, q4.z + q5.w * dx + q3.w * dy};
#pragma unroll
					for (int ch = 0; ch < 3; ch++) {
						S2.x = fmaf(coord[ch], dL_dpixel_coord[ch], S2.x);
						dL_dcoords[ch] = w * dL_dpixel_coord[ch];
						if (is_median) dL_dcoords[ch] += dL_dpixel_mcoord[ch];
						gv[G_VP + ch] = dL_dcoords[ch];
						gv[G_CP + 2 * ch] = dL_dcoords[ch] * dx;      // 1/focal_x applied in backward-preprocess
						gv[G_CP + 2 * ch + 1] = dL_dcoords[ch] * dy;  // 1/focal_y
					}
#pragma unroll
					for (int k = 25; k < GF; k++) gv[k] = 0.f;
				}
				float dL_dt = 0.f;
				if constexpr (DEPTH) {
					dL_dt = c_bt.y;
					if (is_median) dL_dt += dL_dpixel_mt;
				}
				gv[G_T] = dL_dt;
				const float2 g_ray = __fmul2_rn(make_float2(dL_dt, dL_dt), d);
				gv[G_RAYX] = g_ray.x;
				gv[G_RAYY] = g_ray.y;
				float S;
				if constexpr (GEO) {
					S2 = __ffma2_rn(make_float2(q3.x, q3.y), g_nxy, S2);
					const float2 c_n = __fmul2_rn(w2, g_nxy);
					S = fmaf(q3.z, g_nz, S2.x + S2.y);
					gv[G_NRM + 0] = c_n.x;
					gv[G_NRM + 1] = c_n.y;
					gv[G_NRM + 2] = w * g_nz;
				} else {
					S = S2.x + S2.y;
					gv[G_NRM] = gv[G_NRM + 1] = gv[G_NRM + 2] = 0.f;
				}
				const float dL_dopa = T * S - inv * B;
				B = fmaf(w, S, B);
				const float Gm = ok ? G : 0.f;  // the reference skips this pair entirely: every term below carries a factor G

				const float dL_dG = q1.y * dL_dopa;
				const float2 gd = __fmul2_rn(make_float2(Gm, Gm), d);  // (G dx, G dy)
				const float dG_ddelx = -gd.x * q0.z - gd.y * q0.w;
				const float dG_ddely = -gd.y * q1.x - gd.x * q0.w;
				float2 dL_ddel = __fmul2_rn(make_float2(dL_dG, dL_dG), make_float2(dG_ddelx, dG_ddely));
				const float2 ab = __fmul2_rn(dL_ddel, make_float2(ddelx_dx, ddely_dy));
				gv[G_MABS] = abs(ab.x) + abs(ab.y);
				if constexpr (COORD) {
					dL_ddel.x += dL_dcoords[0] * q4.w + dL_dcoords[1] * q5.y + dL_dcoords[2] * q5.w;
					dL_ddel.y += dL_dcoords[0] * q5.x + dL_dcoords[1] * q5.z + dL_dcoords[2] * q3.w;
				}
				if constexpr (DEPTH) dL_ddel = __ffma2_rn(make_float2(dL_dt, dL_dt), make_float2(q1.z, q1.w), dL_ddel);
				gv[G_MX] = dL_ddel.x;  // * 0.5 W in backward-preprocess
				gv[G_MY] = dL_ddel.y;  // * 0.5 H
				const float hG = -0.5f * dL_dG;
				const float2 cxy = __fmul2_rn(__fmul2_rn(make_float2(gd.x, gd.x), d), make_float2(hG, hG));
				gv[G_CONX] = cxy.x;
				gv[G_CONY] = cxy.y;
				gv[G_CONW] = gd.y * dy * hG;
				gv[G_OPA] = Gm * dL_dopa;
				warp_reduce_scatter<GF>(gv, lane);
				float* row = grad_accum + (size_t)lds32(ids_addr + (uint32_t)jj * 4u) * GF;
				if (GF == 16) {
					if ((lane & 1) == 0) atomicAdd(row + (lane >> 1), gv[0]);
				} else {
					if (lane < 25) atomicAdd(row + lane, gv[0]);
				}
			}
		}
	}
}

template <bool COORD, bool DEPTH>
static void launch_variant(const FwdParams& p, GeomView g, BinView b, ImgView img, RenderGradIn gin, float* grad_accum, cudaStream_t s) {
	constexpr int RFQ = COORD ? 6 : 4;
	const size_t smem = (size_t)2 * RFQ * BATCH * sizeof(float4) + (size_t)2 * BATCH * sizeof(int);
	auto kern = render_backward_kernel<COORD, DEPTH>;
	static size_t configured[64] = {};
	ensure_dynamic_smem(kern, smem, configured);
	dim3 grid(p.grid_x, p.row_end - p.row_begin, 1);
	kern<<<grid, NTHREADS, smem, s>>>(img.ranges, b.point_list, g.records, p.W, p.H, p.py_off, p.Hs, p.grid_x, p.row_begin, p.focal_x, p.focal_y, p.background,
	                                 gin.out_alpha, gin.out_normal, img.n_contrib, img.accum_depth, img.accum_coord, img.normal_length,
	                                 gin.d_color, gin.d_coord, gin.d_mcoord, gin.d_depth, gin.d_mdepth, gin.d_alpha, gin.d_normal, grad_accum, img.chunk_base, b.hitmask);
	count_launch();
}

void launch_render_backward(const FwdParams& p, GeomView g, BinView b, ImgView img, RenderGradIn gin, float* grad_accum, cudaStream_t s, bool zero_first) {
	if (zero_first) {
		cudaMemsetAsync(grad_accum, 0, (size_t)p.P * grad_floats(p.coord) * sizeof(float), s);
		count_launch();
	}
	if (p.row_end <= p.row_begin) return;
	if (p.coord && p.depth)
		launch_variant<true, true>(p, g, b, img, gin, grad_accum, s);
	else if (p.coord)
		launch_variant<true, false>(p, g, b, img, gin, grad_accum, s);
	else if (p.depth)
		launch_variant<false, true>(p, g, b, img, gin, grad_accum, s);
	else
		launch_variant<false, false>(p, g, b, img, gin, grad_accum, s);
}

}  // namespace rgs
