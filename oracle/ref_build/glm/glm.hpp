// ORACLE TOOLING ONLY -- not product code, never linked into libb200gs.so.
//
// Minimal stand-in for the GLM headers the reference rasterizer includes
// (`#include <glm/glm.hpp>`, DGR/cuda_rasterizer/forward.h, backward.h,
// rasterizer_impl.cu:23).  third_party/glm is an un-vendored submodule
// (DGR/setup.py:29, DGR/CMakeLists.txt:36) and there is no network, so the
// reference sources are compiled against this shim instead.  It reproduces the
// parts of GLM the reference uses (vec3, vec4, mat3, dot, length, max,
// transpose, operator*), with GLM's column-major storage and GLM's operand
// order inside mat3*mat3 (type_mat3x3.inl: Result[j][i] = a[0][i]*b[j][0] +
// a[1][i]*b[j][1] + a[2][i]*b[j][2]) and dot (x*x + y*y + z*z), so the
// floating-point expression trees nvcc sees are the ones real GLM would give.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define GLM_SHIM_FN __host__ __device__ inline

namespace glm {

struct vec3 {
	float x, y, z;
	GLM_SHIM_FN vec3() : x(0), y(0), z(0) {}
	GLM_SHIM_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
	GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
	GLM_SHIM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	GLM_SHIM_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
	GLM_SHIM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
	float x, y, z, w;
	GLM_SHIM_FN vec4() : x(0), y(0), z(0), w(0) {}
	GLM_SHIM_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

GLM_SHIM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_SHIM_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_SHIM_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_SHIM_FN float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLM_SHIM_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLM_SHIM_FN vec3 max(const vec3& a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }

// Column-major 3x3: m[j] is column j, m[j][i] is row i of column j.
struct mat3 {
	vec3 col[3];
	GLM_SHIM_FN mat3() {}
	GLM_SHIM_FN explicit mat3(float d) {
		col[0] = vec3(d, 0, 0); col[1] = vec3(0, d, 0); col[2] = vec3(0, 0, d);
	}
	GLM_SHIM_FN mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
		col[0] = vec3(x0, y0, z0); col[1] = vec3(x1, y1, z1); col[2] = vec3(x2, y2, z2);
	}
	GLM_SHIM_FN vec3& operator[](int j) { return col[j]; }
	GLM_SHIM_FN const vec3& operator[](int j) const { return col[j]; }
};

GLM_SHIM_FN mat3 transpose(const mat3& m) {
	return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
}

GLM_SHIM_FN mat3 operator*(const mat3& a, const mat3& b) {
	mat3 r;
	for (int j = 0; j < 3; j++)
		for (int i = 0; i < 3; i++)
			r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
	return r;
}

GLM_SHIM_FN mat3 operator*(float s, const mat3& m) {
	mat3 r;
	for (int j = 0; j < 3; j++) r[j] = m[j] * s;
	return r;
}

}  // namespace glm
