// capi_ckpt.cu — restart files.  The reference parses `Restart`, `Restart_Time` and
// `Write_Interval` (main.cu:27-30,47-50) and never uses them; its Fortran predecessor read Data/data.%07d.dat
// (test/UTIL_PRE_SIM.f90:120-150).  Here a checkpoint is the solver's whole time-dependent state as raw fp64 —
// u, v, p with their ghost ring, plus the face velocities in IFX_COMPAT_FULL (they are the convecting
// velocities of the next step) — so a run continued from a file is bit-identical to one that never stopped
// (tests/test_gpu_checkpoint.py).  Bodies are inputs, not state: the caller sets them again.  Slab runs write
// one file per rank (its rows only).
#include "solver.h"

#include <cstdio>
#include <cstring>
#include <vector>

using namespace ifx;

namespace {
struct CkptHeader {
  char magic[8];                 // "IFXCKPT1"
  int abi, compat, nx, ny, j_begin, j_end, rank, nranks, nfields;
  int faces_valid;               // 1: uf, vf are the projected face velocities; 0: they are not state (before the first
                                 // step, after ifx_set_field(u/v), after the bodies moved) and the reader ignores them
  long long step;
  double time;
};
const char kMagic[8] = {'I', 'F', 'X', 'C', 'K', 'P', 'T', '1'};
}  // namespace

extern "C" int ifx_checkpoint_write(ifx_solver* s, const char* path, long long step, double time) {
  if (!s || !path) return IFX_ERR_INVALID;
  const bool full = s->opt.compat == IFX_COMPAT_FULL;
  const ifx_field fields[5] = {IFX_FIELD_U, IFX_FIELD_V, IFX_FIELD_P, IFX_FIELD_UF, IFX_FIELD_VF};
  const int nf = full ? 5 : 3;
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(s, IFX_ERR_IO, std::string("cannot write ") + path);
  CkptHeader h{};
  std::memcpy(h.magic, kMagic, 8);
  h.abi = IFX_ABI_VERSION; h.compat = s->opt.compat; h.nx = s->L.nx; h.ny = s->L.ny;
  h.j_begin = s->L.jb; h.j_end = s->L.je; h.rank = s->opt.rank; h.nranks = s->opt.nranks; h.nfields = nf;
  h.faces_valid = (full && s->faces_valid) ? 1 : 0;
  h.step = step; h.time = time;
  bool ok = std::fwrite(&h, sizeof(h), 1, fp) == 1;
  std::vector<double> buf;
  for (int k = 0; k < nf && ok; k++) {
    const unsigned long long n = ifx_field_size(s, fields[k]);
    buf.resize(n);
    const int rc = ifx_get_field(s, fields[k], buf.data(), n);
    if (rc != IFX_OK) { std::fclose(fp); return rc; }
    const int id = (int)fields[k];
    ok = std::fwrite(&id, sizeof(id), 1, fp) == 1 && std::fwrite(&n, sizeof(n), 1, fp) == 1 &&
         std::fwrite(buf.data(), sizeof(double), n, fp) == n;
  }
  ok = (std::fclose(fp) == 0) && ok;
  return ok ? IFX_OK : fail(s, IFX_ERR_IO, std::string("short write on ") + path);
}

extern "C" int ifx_checkpoint_read(ifx_solver* s, const char* path, long long* step, double* time) {
  if (!s || !path) return IFX_ERR_INVALID;
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(s, IFX_ERR_IO, std::string("cannot open ") + path);
  CkptHeader h{};
  auto bad = [&](const std::string& why) { std::fclose(fp); return fail(s, IFX_ERR_INVALID, std::string(path) + ": " + why); };
  if (std::fread(&h, sizeof(h), 1, fp) != 1 || std::memcmp(h.magic, kMagic, 8) != 0) return bad("not a checkpoint file");
  if (h.nx != s->L.nx || h.ny != s->L.ny) return bad("written for a different grid");
  if (h.j_begin != s->L.jb || h.j_end != s->L.je) return bad("written for a different row slab");
  if (h.compat != s->opt.compat) return bad("written in the other compat mode");
  const int want = s->opt.compat == IFX_COMPAT_FULL ? 5 : 3;
  if (h.nfields != want) return bad("unexpected field count");
  if (h.abi != IFX_ABI_VERSION) return bad("written by another version of the library");
  if (h.rank != s->opt.rank || h.nranks != s->opt.nranks) return bad("written by another rank / for another number of slabs");
  // the whole file is read and validated before any of it is applied: a truncated or foreign file leaves the state alone
  const ifx_field fields[5] = {IFX_FIELD_U, IFX_FIELD_V, IFX_FIELD_P, IFX_FIELD_UF, IFX_FIELD_VF};
  std::vector<std::vector<double>> data(h.nfields);
  for (int k = 0; k < h.nfields; k++) {      // order in the file: u, v, p, then uf, vf
    int id; unsigned long long n;
    if (std::fread(&id, sizeof(id), 1, fp) != 1 || std::fread(&n, sizeof(n), 1, fp) != 1) return bad("truncated");
    if (id != (int)fields[k]) return bad("fields out of order");
    if (n != ifx_field_size(s, fields[k])) return bad("field size mismatch");
    data[k].resize(n);
    if (std::fread(data[k].data(), sizeof(double), n, fp) != n) return bad("truncated");
  }
  std::fclose(fp);
  // bodies set but not classified yet: classify first — ifx_iblank_update invalidates the face velocities (closed
  // faces move with the bodies), which must not happen after the file's faces have been put in place
  if (s->bodies_dirty) {
    const int rc = ifx_iblank_update(s, nullptr);
    if (rc != IFX_OK) return rc;
  }
  // faces after cells (ifx_set_field invalidates the faces when cell velocities are set), and only if they were state
  // when the file was written: otherwise the next predictor rebuilds them from the cells, as the unbroken run does
  for (int k = 0; k < h.nfields; k++) {
    if (k >= 3 && !h.faces_valid) break;
    const int rc = ifx_set_field(s, fields[k], data[k].data(), data[k].size());
    if (rc != IFX_OK) return rc;
  }
  if (step) *step = h.step;
  if (time) *time = h.time;
  return IFX_OK;
}
