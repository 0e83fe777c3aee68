// sph_render.cu -- headless replacement of the reference's GGUI draw calls (main.py:153-161):
//   scene.ambient_light((0.8, 0.8, 0.8)); scene.point_light(pos=(0.5, 1.5, 1.5), color=(1, 1, 1))
//   scene.particles(ps.fluid_particles.pos, radius=ps.particle_radius, per_vertex_color=ps.fluid_particles.rgb)
//   scene.particles(ps.rigid_particles.pos, radius=ps.particle_radius, per_vertex_color=ps.rigid_particles.rgb)
// with the camera of the scene file (cam_pos / cam_look_at / cam_up, main.py:60-62; GGUI's default 45 degree
// vertical field of view).  Particles are splatted as shaded spheres into a depth-tested image that lives in
// device memory: one thread per particle walks the pixels of its projected disc and resolves visibility with a
// 64-bit atomicMin on (depth | colour); no sort, no per-pixel lists.  The image is a caller-owned RGBA8 buffer.
// Mode independent (no solver arithmetic): compiled once.
#include <cmath>

#include "sph_internal.h"

namespace {

struct Cam {
	float px, py, pz;      // position
	float rx, ry, rz;      // right
	float ux, uy, uz;      // up
	float fx, fy, fz;      // forward
	float focal;           // pixels: (height / 2) / tan(fov_y / 2)
	float lx, ly, lz;      // point light position (main.py:154)
	float ambient;
};

__global__ void __launch_bounds__(256) k_render_clear(unsigned long long *zbuf, int n, unsigned int background) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) zbuf[i] = ((unsigned long long)0x7f800000u << 32) | background; // depth = +inf
}

__device__ __forceinline__ unsigned int pack_rgb(float r, float g, float b) {
	unsigned int R = (unsigned int)(fminf(fmaxf(r, 0.0f), 1.0f) * 255.0f + 0.5f);
	unsigned int G = (unsigned int)(fminf(fmaxf(g, 0.0f), 1.0f) * 255.0f + 0.5f);
	unsigned int B = (unsigned int)(fminf(fmaxf(b, 0.0f), 1.0f) * 255.0f + 0.5f);
	return R | (G << 8) | (B << 16) | 0xff000000u;
}

// pos: float4 per particle (xyz); rgb: `rgb_stride` floats per particle (r, g, b first)
__global__ void __launch_bounds__(128)
k_render_splat(const float4 *__restrict__ pos, const float *__restrict__ rgb, int rgb_stride, int n, float radius, Cam cam,
               int width, int height, unsigned long long *zbuf) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 p = pos[i];
	float dx = p.x - cam.px, dy = p.y - cam.py, dz = p.z - cam.pz;
	float xc = dx * cam.rx + dy * cam.ry + dz * cam.rz;
	float yc = dx * cam.ux + dy * cam.uy + dz * cam.uz;
	float zc = dx * cam.fx + dy * cam.fy + dz * cam.fz; // distance along the view direction
	if (zc <= radius) return;                           // behind (or touching) the camera plane
	float inv = cam.focal / zc;
	float sx = 0.5f * width + xc * inv, sy = 0.5f * height - yc * inv;
	float rp = fmaxf(radius * inv, 0.5f);               // at least one pixel
	int x0 = max((int)floorf(sx - rp), 0), x1 = min((int)ceilf(sx + rp), width - 1);
	int y0 = max((int)floorf(sy - rp), 0), y1 = min((int)ceilf(sy + rp), height - 1);
	if (x0 > x1 || y0 > y1) return;
	if ((x1 - x0) > 256 || (y1 - y0) > 256) return;     // a particle in the lens: skip rather than stall a thread
	float cr = rgb[(size_t)i * rgb_stride], cg = rgb[(size_t)i * rgb_stride + 1], cb = rgb[(size_t)i * rgb_stride + 2];
	float inv_rp = 1.0f / rp;
	for (int y = y0; y <= y1; ++y) {
		for (int x = x0; x <= x1; ++x) {
			float u = ((float)x + 0.5f - sx) * inv_rp, v = ((float)y + 0.5f - sy) * inv_rp;
			float d2 = u * u + v * v;
			if (d2 > 1.0f && rp > 0.5f) continue;
			float nz = sqrtf(fmaxf(1.0f - d2, 0.0f)); // sphere normal in camera space: (u, -v, -nz)
			float depth = zc - nz * radius;
			// surface point and normal in world space
			float nxw = u * cam.rx - v * cam.ux - nz * cam.fx;
			float nyw = u * cam.ry - v * cam.uy - nz * cam.fy;
			float nzw = u * cam.rz - v * cam.uz - nz * cam.fz;
			float wx = p.x + nxw * radius, wy = p.y + nyw * radius, wz = p.z + nzw * radius;
			float lx = cam.lx - wx, ly = cam.ly - wy, lz = cam.lz - wz;
			float ll = rsqrtf(fmaxf(lx * lx + ly * ly + lz * lz, 1e-20f));
			float diff = fmaxf((nxw * lx + nyw * ly + nzw * lz) * ll, 0.0f);
			float shade = fminf(cam.ambient + diff, 1.0f);
			unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | pack_rgb(cr * shade, cg * shade, cb * shade);
			atomicMin(&zbuf[(size_t)y * width + x], key); // positive floats order like their bit patterns
		}
	}
}

__global__ void __launch_bounds__(256) k_render_resolve(const unsigned long long *__restrict__ zbuf, int n, unsigned int *__restrict__ rgba,
                                                        float *__restrict__ depth) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned long long k = zbuf[i];
	rgba[i] = (unsigned int)(k & 0xffffffffu);
	if (depth) depth[i] = __uint_as_float((unsigned int)(k >> 32));
}

static void normalize(double v[3]) {
	double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	if (l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; }
}

} // namespace

extern "C" int sph_render(SphHandle *h, const SphCamera *camera, int width, int height, int what, const void *dev_fluid_rgb,
                          int fluid_rgb_stride, const void *dev_rigid_rgb, int rigid_rgb_stride, void *dev_rgba8,
                          void *dev_depth, void *stream) {
	if (!h || !camera || !dev_rgba8) return SPH_EINVAL;
	if (width <= 0 || height <= 0 || (long long)width * height > (1LL << 28))
		return sph_fail(h, SPH_EINVAL, "sph_render: bad image size %d x %d", width, height);
	if (!h->pos) return sph_fail(h, SPH_ENOTBOUND, "sph_render: fluid positions are not bound");
	cudaStream_t st = (cudaStream_t)stream;
	SPH_CUDA_CHECK(h, cudaSetDevice(h->device));
	size_t npix = (size_t)width * (size_t)height;
	if (h->render_cap < npix) {
		cudaFree(h->render_zbuf);
		h->render_zbuf = nullptr;
		h->render_cap = 0;
		SPH_CUDA_CHECK(h, cudaMalloc((void **)&h->render_zbuf, sizeof(unsigned long long) * npix));
		h->render_cap = npix;
	}
	// camera frame (right-handed, like GGUI): forward = look_at - pos, right = forward x up, up' = right x forward
	double f[3] = {camera->look_at[0] - camera->pos[0], camera->look_at[1] - camera->pos[1], camera->look_at[2] - camera->pos[2]};
	normalize(f);
	double up[3] = {camera->up[0], camera->up[1], camera->up[2]};
	double r[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
	normalize(r);
	double u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
	double fov = camera->fov_y_deg > 0 ? camera->fov_y_deg : 45.0;
	Cam c;
	c.px = (float)camera->pos[0]; c.py = (float)camera->pos[1]; c.pz = (float)camera->pos[2];
	c.rx = (float)r[0]; c.ry = (float)r[1]; c.rz = (float)r[2];
	c.ux = (float)u[0]; c.uy = (float)u[1]; c.uz = (float)u[2];
	c.fx = (float)f[0]; c.fy = (float)f[1]; c.fz = (float)f[2];
	c.focal = (float)(0.5 * height / tan(0.5 * fov * 3.14159265358979323846 / 180.0));
	c.lx = (float)camera->light_pos[0]; c.ly = (float)camera->light_pos[1]; c.lz = (float)camera->light_pos[2];
	c.ambient = (float)camera->ambient;
	unsigned int bg = 0xff000000u | ((unsigned int)camera->background[0]) | ((unsigned int)camera->background[1] << 8) |
	                  ((unsigned int)camera->background[2] << 16);
	int nb = (int)((npix + 255) / 256);
	k_render_clear<<<nb, 256, 0, st>>>(h->render_zbuf, (int)npix, bg);
	float radius = (float)h->cfg.particle_radius;
	int nf = h->c.N_owned;
	if ((what & SPH_RENDER_FLUID) && nf > 0 && dev_fluid_rgb)
		k_render_splat<<<(nf + 127) / 128, 128, 0, st>>>(h->pos, (const float *)dev_fluid_rgb, fluid_rgb_stride, nf, radius, c, width,
		                                                height, h->render_zbuf);
	if ((what & SPH_RENDER_RIGID) && h->c.Nr > 0 && h->rpos && dev_rigid_rgb)
		k_render_splat<<<(h->c.Nr + 127) / 128, 128, 0, st>>>(h->rpos, (const float *)dev_rigid_rgb, rigid_rgb_stride, h->c.Nr, radius,
		                                                      c, width, height, h->render_zbuf);
	k_render_resolve<<<nb, 256, 0, st>>>(h->render_zbuf, (int)npix, (unsigned int *)dev_rgba8, (float *)dev_depth);
	h->launches += 4;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return sph_fail(h, SPH_ECUDA, "sph_render: %s", cudaGetErrorString(e));
	return SPH_OK;
}
