// STAND-IN for muu/matrix.h (see vector.h).  Column-major storage, right-handed, clip-space depth in [0,1]
// (UNVERIFIED against muu): the same conventions rt_b200/camera.py restates.  TEST INFRASTRUCTURE.
#pragma once
#include "vector.h"

namespace muu
{
	template <typename T, std::size_t R, std::size_t C>
	struct matrix;

	template <typename T>
	struct matrix<T, 3, 3>
	{
		vector<T, 3> m[3]{}; // columns
		struct constants_t
		{
			static constexpr matrix identity_() noexcept
			{
				matrix r;
				r.m[0] = { 1, 0, 0 }; r.m[1] = { 0, 1, 0 }; r.m[2] = { 0, 0, 1 };
				return r;
			}
		};
		struct constants
		{
			static constexpr matrix identity = [] {
				matrix r;
				r.m[0] = { T{ 1 }, T{}, T{} }; r.m[1] = { T{}, T{ 1 }, T{} }; r.m[2] = { T{}, T{}, T{ 1 } };
				return r;
			}();
		};
		constexpr T operator()(std::size_t r, std::size_t c) const noexcept { return m[c][r]; }
		[[nodiscard]] constexpr vector<T, 3> transform_direction(vector<T, 3> v) const noexcept
		{
			return m[0] * v.x + m[1] * v.y + m[2] * v.z;
		}
		friend constexpr matrix operator*(const matrix& a, const matrix& b) noexcept
		{
			matrix r;
			for (int c = 0; c < 3; c++)
				r.m[c] = a.transform_direction(b.m[c]);
			return r;
		}
		// Gram-Schmidt on the columns
		[[nodiscard]] static constexpr matrix orthonormalize(const matrix& in) noexcept
		{
			using v3 = vector<T, 3>;
			matrix r;
			r.m[0] = v3::normalize(in.m[0]);
			r.m[1] = v3::normalize(in.m[1] - r.m[0] * v3::dot(r.m[0], in.m[1]));
			r.m[2] = v3::normalize(in.m[2] - r.m[0] * v3::dot(r.m[0], in.m[2]) - r.m[1] * v3::dot(r.m[1], in.m[2]));
			return r;
		}
		[[nodiscard]] static constexpr matrix from_axis_angle(vector<T, 3> axis, T angle) noexcept
		{
			using v3 = vector<T, 3>;
			const v3 a = v3::normalize(axis);
			const T c = __builtin_cosf(angle), s = __builtin_sinf(angle), t = T{ 1 } - c;
			matrix r;
			r.m[0] = { t * a.x * a.x + c, t * a.x * a.y + s * a.z, t * a.x * a.z - s * a.y };
			r.m[1] = { t * a.x * a.y - s * a.z, t * a.y * a.y + c, t * a.y * a.z + s * a.x };
			r.m[2] = { t * a.x * a.z + s * a.y, t * a.y * a.z - s * a.x, t * a.z * a.z + c };
			return r;
		}
		// a rotation whose forward (-Z) axis is `dir` (camera.hpp:116-119); same construction as rt_b200/camera.py
		[[nodiscard]] static constexpr matrix from_3d_direction(vector<T, 3> dir) noexcept
		{
			using v3 = vector<T, 3>;
			const v3 f = v3::normalize(dir);
			const v3 back = -f;
			v3 up = v3::constants::up;
			if (__builtin_fabsf(v3::dot(f, up)) >= T(0.9999))
				up = v3{ T{}, T{}, f.y < T{} ? T{ 1 } : T{ -1 } };
			const v3 right = v3::normalize(v3::cross(up, back));
			const v3 up2 = v3::cross(back, right);
			matrix r;
			r.m[0] = right; r.m[1] = up2; r.m[2] = back;
			return r;
		}
	};

	template <typename T>
	[[nodiscard]] constexpr matrix<T, 3, 3> orthonormalize(const matrix<T, 3, 3>& in) noexcept
	{
		return matrix<T, 3, 3>::orthonormalize(in);
	}
	template <typename T, std::size_t R, std::size_t C>
	[[nodiscard]] constexpr bool infinity_or_nan(const matrix<T, R, C>&) noexcept
	{
		return false;
	}

	template <typename T>
	struct matrix<T, 4, 4>
	{
		vector<T, 4> m[4]{}; // columns
		constexpr T operator()(std::size_t r, std::size_t c) const noexcept
		{
			const vector<T, 4>& col = m[c];
			return r == 0 ? col.x : (r == 1 ? col.y : (r == 2 ? col.z : col.w));
		}
		constexpr void set(std::size_t r, std::size_t c, T v) noexcept
		{
			vector<T, 4>& col = m[c];
			(r == 0 ? col.x : (r == 1 ? col.y : (r == 2 ? col.z : col.w))) = v;
		}
		[[nodiscard]] static constexpr matrix identity() noexcept
		{
			matrix r;
			for (int i = 0; i < 4; i++) r.set(i, i, T{ 1 });
			return r;
		}
		[[nodiscard]] static constexpr matrix from_translation(vector<T, 3> t) noexcept
		{
			matrix r = identity();
			r.m[3] = { t.x, t.y, t.z, T{ 1 } };
			return r;
		}
		[[nodiscard]] static constexpr matrix from_3d_rotation(const matrix<T, 3, 3>& rot) noexcept
		{
			matrix r = identity();
			for (int c = 0; c < 3; c++) r.m[c] = { rot.m[c].x, rot.m[c].y, rot.m[c].z, T{} };
			return r;
		}
		// right-handed, depth 0..1 (near -> 0, far -> 1), vertical field of view, aspect from the size
		[[nodiscard]] static constexpr matrix perspective_projection(T vfov, vector<T, 2> size, T near_, T far_) noexcept
		{
			const double f = 1.0 / __builtin_tan(static_cast<double>(vfov) / 2.0);
			const double aspect = static_cast<double>(size.x) / static_cast<double>(size.y);
			matrix r;
			r.set(0, 0, static_cast<T>(f / aspect));
			r.set(1, 1, static_cast<T>(f));
			r.set(2, 2, static_cast<T>(static_cast<double>(far_) / (static_cast<double>(near_) - far_)));
			r.set(2, 3, static_cast<T>(static_cast<double>(near_) * far_ / (static_cast<double>(near_) - far_)));
			r.set(3, 2, T{ -1 });
			return r;
		}
		friend constexpr vector<T, 4> operator*(const matrix& a, vector<T, 4> v) noexcept
		{
			return a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z + a.m[3] * v.w;
		}
		friend constexpr matrix operator*(const matrix& a, const matrix& b) noexcept
		{
			matrix r;
			for (int c = 0; c < 4; c++) r.m[c] = a * b.m[c];
			return r;
		}
		// general inverse by Gauss-Jordan in double (the matrix is *input data* of the hot path; SPEC S7 starts from it)
		[[nodiscard]] static constexpr matrix invert(const matrix& in) noexcept
		{
			double a[4][8] = {};
			for (int r = 0; r < 4; r++)
				for (int c = 0; c < 4; c++)
				{
					a[r][c] = static_cast<double>(in(r, c));
					a[r][4 + c] = r == c ? 1.0 : 0.0;
				}
			for (int col = 0; col < 4; col++)
			{
				int piv = col;
				for (int r = col + 1; r < 4; r++)
					if (__builtin_fabs(a[r][col]) > __builtin_fabs(a[piv][col])) piv = r;
				for (int c = 0; c < 8; c++) { const double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
				const double d = a[col][col];
				for (int c = 0; c < 8; c++) a[col][c] /= d;
				for (int r = 0; r < 4; r++)
					if (r != col)
					{
						const double f = a[r][col];
						for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
					}
			}
			matrix out;
			for (int r = 0; r < 4; r++)
				for (int c = 0; c < 4; c++) out.set(r, c, static_cast<T>(a[r][4 + c]));
			return out;
		}
		// SPEC S7: M * (p, 1), perspective divide by w through one reciprocal
		[[nodiscard]] vector<T, 3> transform_position(vector<T, 3> p) const noexcept
		{
			T h[4];
			for (int r = 0; r < 4; r++)
				h[r] = shim_fma((*this)(r, 2), p.z, shim_fma((*this)(r, 1), p.y, shim_fma((*this)(r, 0), p.x, (*this)(r, 3))));
			const T inv_w = T{ 1 } / h[3];
			if (p.z == T{ 1 })
				shim::on_primary_done();
			return { h[0] * inv_w, h[1] * inv_w, h[2] * inv_w };
		}
	};
}
