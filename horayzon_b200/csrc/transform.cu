// transform.cu -- coordinate preparation on the device (scope row "next 3").
//
// Replaces the single-threaded Cython loops that turn a lon/lat/elevation grid
// into the inputs of the horizon / shadow path (all file:line relative to the
// reference's horayzon/ directory):
//   _lonlat2ecef_1d       transform.pyx:60-103     geodetic -> ECEF (fp64)
//   _ecef2enu_1d          transform.pyx:152-189    ECEF -> local tangent plane (fp64 in, fp32 out)
//   _ecef2enu_vector_1d   transform.pyx:231-261    vectors ECEF -> ENU (fp32)
//   _wgs2swiss_1d         transform.pyx:306-344    WGS84 -> LV95 (approximate formulas)
//   _swiss2wgs_1d         transform.pyx:390-432    LV95 -> WGS84
//   rotation_matrix_glob2loc  transform.pyx:490-530    rows east = north x norm, north, norm
//   _surf_norm_1d         direction.pyx:48-70      ellipsoid normal (n-vector)
//   _north_dir_1d         direction.pyx:125-178    unit vector towards the North pole
// One thread per element, the reference's operation order in double; sin/cos are
// CUDA's double-precision routines (<= 2 ulp; the reference itself is compiled
// with -ffast-math, so bit identity is not defined -- tests state the tolerance).
//
// hzb_prep_enu_dev is the additive fused form: lon (nx), lat (ny), elevation
// (ny x nx) -> interleaved ENU vertex buffer + ENU normals / north vectors of the
// inner domain, all in HBM, with the same device functions as the step-by-step
// path (so both give identical bits).
#include "hzb_common.cuh"
#include <math.h>

namespace hzb {
namespace {

struct Ellps { double a_or_r, b2_a2, e_2, np_z; int sphere; };

// transform.pyx:76-101 and direction.pyx:141-154
Ellps make_ellps(int ellps) {
    Ellps e;
    if (ellps == 0) { e.sphere = 1; e.a_or_r = 6370997.0; e.b2_a2 = 1.0; e.e_2 = 0.0; e.np_z = 6370997.0; return e; }
    const double a = 6378137.0;
    const double f = ellps == 1 ? (1.0 / 298.257222101) : (1.0 / 298.257223563);   // GRS80 / WGS84
    const double b = a * (1.0 - f);
    e.sphere = 0; e.a_or_r = a; e.b2_a2 = (b * b) / (a * a); e.e_2 = 1.0 - (b * b) / (a * a); e.np_z = b;
    return e;
}

__device__ __forceinline__ double deg2rad_dd(double a) { return a * (M_PI / 180.0); }   // transform.pyx:537-542

__device__ __forceinline__ void lonlat2ecef_one(const Ellps& E, double lon, double lat, float h, double& x, double& y, double& z) {
    const double sl = sin(deg2rad_dd(lat)), cl = cos(deg2rad_dd(lat));
    const double so = sin(deg2rad_dd(lon)), co = cos(deg2rad_dd(lon));
    if (E.sphere) {                                   // :78-85
        const double r = E.a_or_r + (double)h;
        x = r * cl * co; y = r * cl * so; z = r * sl;
    } else {                                          // :96-101
        const double n = E.a_or_r / sqrt(1.0 - E.e_2 * (sl * sl));
        x = (n + (double)h) * cl * co; y = (n + (double)h) * cl * so; z = (E.b2_a2 * n + (double)h) * sl;
    }
}

struct EnuFrame { double x0, y0, z0, sin_lon, cos_lon, sin_lat, cos_lat; };   // TransformerEcef2enu + :170-173

__device__ __forceinline__ void ecef2enu_one(const EnuFrame& F, double x, double y, double z, float& xe, float& ye, float& ze) {
    const double dx = x - F.x0, dy = y - F.y0, dz = z - F.z0;                 // :177-187
    xe = (float)(-F.sin_lon * dx + F.cos_lon * dy);
    ye = (float)(-F.sin_lat * F.cos_lon * dx - F.sin_lat * F.sin_lon * dy + F.cos_lat * dz);
    ze = (float)(F.cos_lat * F.cos_lon * dx + F.cos_lat * F.sin_lon * dy + F.sin_lat * dz);
}
__device__ __forceinline__ void ecef2enu_vec_one(const EnuFrame& F, float vx, float vy, float vz, float& xe, float& ye, float& ze) {
    xe = (float)(-F.sin_lon * (double)vx + F.cos_lon * (double)vy);           // :251-259
    ye = (float)(-F.sin_lat * F.cos_lon * (double)vx - F.sin_lat * F.sin_lon * (double)vy + F.cos_lat * (double)vz);
    ze = (float)(F.cos_lat * F.cos_lon * (double)vx + F.cos_lat * F.sin_lon * (double)vy + F.sin_lat * (double)vz);
}
__device__ __forceinline__ void surf_norm_one(double lon, double lat, float& nx, float& ny, float& nz) {
    const double so = sin(deg2rad_dd(lon)), co = cos(deg2rad_dd(lon));        // direction.pyx:61-68
    const double sl = sin(deg2rad_dd(lat)), cl = cos(deg2rad_dd(lat));
    nx = (float)(cl * co); ny = (float)(cl * so); nz = (float)sl;
}
__device__ __forceinline__ void north_dir_one(double np_z, double x, double y, double z, float nx, float ny, float nz,
                                              float& ox, float& oy, float& oz) {
    const double vx = 0.0 - x, vy = 0.0 - y, vz = np_z - z;                   // direction.pyx:159-176
    const double dp = (vx * (double)nx) + (vy * (double)ny) + (vz * (double)nz);
    const double px = vx - dp * (double)nx, py = vy - dp * (double)ny, pz = vz - dp * (double)nz;
    const double nrm = sqrt(px * px + py * py + pz * pz);
    ox = (float)(px / nrm); oy = (float)(py / nrm); oz = (float)(pz / nrm);
}

__global__ void k_lonlat2ecef(Ellps E, const double* __restrict__ lon, const double* __restrict__ lat, const float* __restrict__ h,
                              long long n, double* __restrict__ x, double* __restrict__ y, double* __restrict__ z) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) lonlat2ecef_one(E, lon[i], lat[i], h[i], x[i], y[i], z[i]);
}
__global__ void k_ecef2enu(EnuFrame F, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                           long long n, float* __restrict__ xe, float* __restrict__ ye, float* __restrict__ ze) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) ecef2enu_one(F, x[i], y[i], z[i], xe[i], ye[i], ze[i]);
}
__global__ void k_ecef2enu_vec(EnuFrame F, const float* __restrict__ v, long long n, float* __restrict__ o) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) ecef2enu_vec_one(F, v[3 * i], v[3 * i + 1], v[3 * i + 2], o[3 * i], o[3 * i + 1], o[3 * i + 2]);
}
__global__ void k_surf_norm(const double* __restrict__ lon, const double* __restrict__ lat, long long n, float* __restrict__ o) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) surf_norm_one(lon[i], lat[i], o[3 * i], o[3 * i + 1], o[3 * i + 2]);
}
__global__ void k_north_dir(double np_z, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                            const float* __restrict__ nrm, long long n, float* __restrict__ o) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) north_dir_one(np_z, x[i], y[i], z[i], nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], o[3 * i], o[3 * i + 1], o[3 * i + 2]);
}
__global__ void k_wgs2swiss(const double* __restrict__ lon, const double* __restrict__ lat, const float* __restrict__ h, long long n,
                            double* __restrict__ e, double* __restrict__ nn, float* __restrict__ hc) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lo = ((lon[i] * 3600.0) - 26782.5) / 10000.0, la = ((lat[i] * 3600.0) - 169028.66) / 10000.0;   // :325-342
    e[i] = 2600072.37 + 211455.93 * lo - 10938.51 * lo * la - 0.36 * lo * (la * la) - 44.54 * (lo * lo * lo);
    nn[i] = 1200147.07 + 308807.95 * la + 3745.25 * (lo * lo) + 76.63 * (la * la) - 194.56 * (lo * lo) * la + 119.79 * (la * la * la);
    hc[i] = (float)((double)h[i] - 49.55 + 2.73 * lo + 6.94 * la);
}
__global__ void k_swiss2wgs(const double* __restrict__ e, const double* __restrict__ nn, const float* __restrict__ hc, long long n,
                            double* __restrict__ lon, double* __restrict__ lat, float* __restrict__ h) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ep = (e[i] - 2600000.0) / 1000000.0, np_ = (nn[i] - 1200000.0) / 1000000.0;                      // :409-430
    double lo = 2.6779094 + 4.728982 * ep + 0.791484 * ep * np_ + 0.1306 * ep * (np_ * np_) - 0.0436 * (ep * ep * ep);
    double la = 16.9023892 + 3.238272 * np_ - 0.270978 * (ep * ep) - 0.002528 * (np_ * np_) - 0.0447 * (ep * ep) * np_ - 0.0140 * (np_ * np_ * np_);
    h[i] = (float)((double)hc[i] + 49.55 - 12.60 * ep - 22.64 * np_);
    lon[i] = lo * (100.0 / 36.); lat[i] = la * (100.0 / 36.);
}
// rotation_matrix_glob2loc (transform.pyx:516-530): out [(ny+2)][(nx+2)][3][3], NaN rim
__global__ void k_rotmat(const float* __restrict__ north, const float* __restrict__ norm, int ny, int nx, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)(ny + 2) * (nx + 2);
    if (i >= total) return;
    const int r = (int)(i / (nx + 2)), c = (int)(i - (long long)r * (nx + 2));
    float* o = out + 9 * i;
    if (r == 0 || c == 0 || r == ny + 1 || c == nx + 1) {
        const float nanv = __int_as_float(0x7fc00000);
        for (int k = 0; k < 9; ++k) o[k] = nanv;
        return;
    }
    const long long s = (long long)(r - 1) * nx + (c - 1);
    const float ax = north[3 * s], ay = north[3 * s + 1], az = north[3 * s + 2];
    const float bx = norm[3 * s], by = norm[3 * s + 1], bz = norm[3 * s + 2];
    o[0] = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));     // np.cross(north, norm) in float32
    o[1] = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    o[2] = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
    o[3] = ax; o[4] = ay; o[5] = az; o[6] = bx; o[7] = by; o[8] = bz;
}

// Fused: lon[nx], lat[ny], elev[ny][nx] -> vert_grid[ny][nx][3] (ENU, fp32) and, for the inner domain
// [off0, off0+in0) x [off1, off1+in1), vec_norm / vec_north in ENU.
__global__ void k_prep_enu(Ellps E, EnuFrame F, const double* __restrict__ lon, const double* __restrict__ lat,
                           const float* __restrict__ elev, int ny, int nx, int off0, int off1, int in0, int in1,
                           float* __restrict__ vert_grid, float* __restrict__ vec_norm, float* __restrict__ vec_north) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)ny * nx) return;
    const int r = (int)(i / nx), c = (int)(i - (long long)r * nx);
    double x, y, z;
    lonlat2ecef_one(E, lon[c], lat[r], elev[i], x, y, z);
    ecef2enu_one(F, x, y, z, vert_grid[3 * i], vert_grid[3 * i + 1], vert_grid[3 * i + 2]);
    const int ri = r - off0, ci = c - off1;
    if (vec_norm && ri >= 0 && ri < in0 && ci >= 0 && ci < in1) {
        const long long o = (long long)ri * in1 + ci;
        float nx_, ny_, nz_, tx, ty, tz;
        surf_norm_one(lon[c], lat[r], nx_, ny_, nz_);
        north_dir_one(E.np_z, x, y, z, nx_, ny_, nz_, tx, ty, tz);
        ecef2enu_vec_one(F, nx_, ny_, nz_, vec_norm[3 * o], vec_norm[3 * o + 1], vec_norm[3 * o + 2]);
        ecef2enu_vec_one(F, tx, ty, tz, vec_north[3 * o], vec_north[3 * o + 1], vec_north[3 * o + 2]);
    }
}

inline unsigned int nblk(long long n) { return (unsigned int)((n + 255) / 256); }

}  // namespace

// The origin frame: x/y/z_ecef_or come from the caller (TransformerEcef2enu attributes); trig in double
// on the host like transform.pyx:170-173.
static EnuFrame make_frame_enu(double x0, double y0, double z0, double lon_or, double lat_or) {
    EnuFrame F;
    F.x0 = x0; F.y0 = y0; F.z0 = z0;
    const double lo = lon_or * (M_PI / 180.0), la = lat_or * (M_PI / 180.0);
    F.sin_lon = sin(lo); F.cos_lon = cos(lo); F.sin_lat = sin(la); F.cos_lat = cos(la);
    return F;
}

int launch_lonlat2ecef(int ellps, const double* lon, const double* lat, const float* h, long long n, double* x, double* y, double* z, cudaStream_t st) {
    if (n <= 0) return 0;
    k_lonlat2ecef<<<nblk(n), 256, 0, st>>>(make_ellps(ellps), lon, lat, h, n, x, y, z);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_ecef2enu(const double* x, const double* y, const double* z, long long n, double x0, double y0, double z0, double lon_or, double lat_or,
                    float* xe, float* ye, float* ze, cudaStream_t st) {
    if (n <= 0) return 0;
    k_ecef2enu<<<nblk(n), 256, 0, st>>>(make_frame_enu(x0, y0, z0, lon_or, lat_or), x, y, z, n, xe, ye, ze);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_ecef2enu_vector(const float* v, long long n, double lon_or, double lat_or, float* o, cudaStream_t st) {
    if (n <= 0) return 0;
    k_ecef2enu_vec<<<nblk(n), 256, 0, st>>>(make_frame_enu(0, 0, 0, lon_or, lat_or), v, n, o);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_surf_norm(const double* lon, const double* lat, long long n, float* o, cudaStream_t st) {
    if (n <= 0) return 0;
    k_surf_norm<<<nblk(n), 256, 0, st>>>(lon, lat, n, o);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_north_dir(int ellps, const double* x, const double* y, const double* z, const float* nrm, long long n, float* o, cudaStream_t st) {
    if (n <= 0) return 0;
    k_north_dir<<<nblk(n), 256, 0, st>>>(make_ellps(ellps).np_z, x, y, z, nrm, n, o);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_wgs2swiss(const double* lon, const double* lat, const float* h, long long n, double* e, double* nn, float* hc, cudaStream_t st) {
    if (n <= 0) return 0;
    k_wgs2swiss<<<nblk(n), 256, 0, st>>>(lon, lat, h, n, e, nn, hc);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_swiss2wgs(const double* e, const double* nn, const float* hc, long long n, double* lon, double* lat, float* h, cudaStream_t st) {
    if (n <= 0) return 0;
    k_swiss2wgs<<<nblk(n), 256, 0, st>>>(e, nn, hc, n, lon, lat, h);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_rotmat(const float* north, const float* norm, int ny, int nx, float* out, cudaStream_t st) {
    k_rotmat<<<nblk((long long)(ny + 2) * (nx + 2)), 256, 0, st>>>(north, norm, ny, nx, out);
    HZB_CUDA(cudaGetLastError()); return 0;
}
int launch_prep_enu(int ellps, const double* lon, const double* lat, const float* elev, int ny, int nx, double x0, double y0, double z0,
                    double lon_or, double lat_or, int off0, int off1, int in0, int in1, float* vert_grid, float* vec_norm,
                    float* vec_north, cudaStream_t st) {
    if (ny <= 0 || nx <= 0) return 0;
    k_prep_enu<<<nblk((long long)ny * nx), 256, 0, st>>>(make_ellps(ellps), make_frame_enu(x0, y0, z0, lon_or, lat_or), lon, lat, elev,
                                                         ny, nx, off0, off1, in0, in1, vert_grid, vec_norm, vec_north);
    HZB_CUDA(cudaGetLastError()); return 0;
}

}  // namespace hzb
