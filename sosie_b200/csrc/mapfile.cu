// mapfile.cu -- raw-binary twin of the mapping file.
//
// The reference keeps the bilinear mapping in a NetCDF-4 file, `sosie_mapping_<pattern>.nc`, written by
// P2D_MAPPING_AB (src/io_ezcdf.f90:2195-2274) and read by RD_MAPPING_AB (:2280-2304).  libnetcdf/HDF5 do not exist
// in this environment and the Fortran shim keeps using the reference's own NetCDF routines, so this file defines the
// documented binary twin: the same variables, in the same order, types and (column-major) layout, behind a fixed
// 64-byte header.  A CPU-built and a GPU-built mapping are interchangeable through either container.
//
//   offset  0  char[8]  "SOSIEMAP"
//           8  int32    version (1)
//          12  int32    nx          (dimension 'x')
//          16  int32    ny          (dimension 'y')
//          20  int32    has_dist_np (the optional d2np argument of P2D_MAPPING_AB)
//          24  float64  vflag       (the missing-value attribute written on metrics / alphabeta)
//          32  ..63     zero
//   then   lon(nx,ny) f64 | lat(nx,ny) f64 | metrics(nx,ny,3) i32 | alphabeta(nx,ny,2) f64 | iproblem(nx,ny) i32
//          (NF90_INT in the NetCDF file although the Fortran array is INTEGER(2)) | dist_np(nx,ny) f32 (if present)
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/sosie_b200.h"

namespace {
struct Header {
  char magic[8];
  int32_t version, nx, ny, has_dnp;
  double vflag;
  char pad[32];
};
static_assert(sizeof(Header) == 64, "mapping header must be 64 bytes");
}  // namespace

extern "C" {

int sosie_mapping_write_bin(const char* path, int32_t nx, int32_t ny, const double* lon, const double* lat, const int32_t* metrics,
                            const double* alphabeta, const int16_t* id_problem, const float* dist_np, double vflag) {
  if (!path || nx < 1 || ny < 1 || !lon || !lat || !metrics || !alphabeta || !id_problem) return SOSIE_ERR_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return SOSIE_ERR_ARG;
  Header h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "SOSIEMAP", 8);
  h.version = 1; h.nx = nx; h.ny = ny; h.has_dnp = dist_np ? 1 : 0; h.vflag = vflag;
  const size_t n = (size_t)nx * ny;
  std::vector<int32_t> ipb(n);
  for (size_t k = 0; k < n; k++) ipb[k] = (int32_t)id_problem[k];
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(lon, 8, n, f) == n && fwrite(lat, 8, n, f) == n &&
            fwrite(metrics, 4, 3 * n, f) == 3 * n && fwrite(alphabeta, 8, 2 * n, f) == 2 * n && fwrite(ipb.data(), 4, n, f) == n;
  if (ok && dist_np) ok = fwrite(dist_np, 4, n, f) == n;
  ok = (fclose(f) == 0) && ok;
  return ok ? SOSIE_OK : SOSIE_ERR_ARG;
}

int sosie_mapping_read_header_bin(const char* path, int32_t* nx, int32_t* ny, int32_t* has_dist_np) {
  if (!path || !nx || !ny) return SOSIE_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return SOSIE_ERR_ARG;
  Header h;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "SOSIEMAP", 8) == 0 && h.version == 1 && h.nx > 0 && h.ny > 0;
  fclose(f);
  if (!ok) return SOSIE_ERR_ARG;
  *nx = h.nx; *ny = h.ny;
  if (has_dist_np) *has_dist_np = h.has_dnp;
  return SOSIE_OK;
}

/* RD_MAPPING_AB twin: metrics, alphabeta, iproblem (any of lon / lat / dist_np may be NULL: skipped) */
int sosie_mapping_read_bin(const char* path, int32_t nx, int32_t ny, double* lon, double* lat, int32_t* metrics, double* alphabeta,
                           int16_t* id_problem, float* dist_np) {
  if (!path || !metrics || !alphabeta || !id_problem) return SOSIE_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return SOSIE_ERR_ARG;
  Header h;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "SOSIEMAP", 8) == 0 && h.version == 1 && h.nx == nx && h.ny == ny;
  const size_t n = (size_t)nx * ny;
  auto rd = [&](void* dst, size_t esz, size_t cnt) {
    if (!ok) return;
    if (dst) ok = fread(dst, esz, cnt, f) == cnt;
    else ok = fseek(f, (long)(esz * cnt), SEEK_CUR) == 0;
  };
  rd(lon, 8, n); rd(lat, 8, n); rd(metrics, 4, 3 * n); rd(alphabeta, 8, 2 * n);
  if (ok) {
    std::vector<int32_t> ipb(n);
    ok = fread(ipb.data(), 4, n, f) == n;
    if (ok) for (size_t k = 0; k < n; k++) id_problem[k] = (int16_t)ipb[k];
  }
  if (ok && dist_np) {
    if (h.has_dnp) ok = fread(dist_np, 4, n, f) == n;
    else for (size_t k = 0; k < n; k++) dist_np[k] = 0.f;
  }
  fclose(f);
  return ok ? SOSIE_OK : SOSIE_ERR_ARG;
}

}  // extern "C"
