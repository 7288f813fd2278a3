// mrg_restart.cpp -- see mrg_restart.h.
#include "mrg_restart.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
uint64_t kMaxSub = 2147483639ull;   // gfortran's default maximum subrecord length, 2^31 - 9 bytes (-fmax-subrecord-length)

// writes `bytes` of a record body that is being emitted as subrecords; `pos` = bytes already written into the current
// subrecord, `sublen` = its length
struct SubWriter {
  FILE* f;
  uint64_t total, done = 0, sub_done = 0, sub_len = 0;
  bool first = true, ok = true;
  void open_sub() {
    const uint64_t left = total - done;
    sub_len = left > kMaxSub ? kMaxSub : left;
    const bool more = left > sub_len;
    const int32_t m = more ? -(int32_t)sub_len : (int32_t)sub_len;       // leading marker: negative = continued in the next subrecord
    ok = ok && fwrite(&m, 4, 1, f) == 1;
    sub_done = 0;
  }
  void close_sub() {
    const int32_t m = first ? (int32_t)sub_len : -(int32_t)sub_len;      // trailing marker: negative = has a preceding subrecord
    ok = ok && fwrite(&m, 4, 1, f) == 1;
    first = false;
  }
  void write(const void* p, uint64_t n) {
    const char* c = (const char*)p;
    while (n > 0 && ok) {
      if (sub_done == sub_len && done < total) { if (done > 0) close_sub(); open_sub(); }
      const uint64_t k = (sub_len - sub_done) < n ? (sub_len - sub_done) : n;
      ok = ok && fwrite(c, 1, (size_t)k, f) == (size_t)k;
      c += k; n -= k; sub_done += k; done += k;
    }
  }
  void finish() {
    if (total == 0) { open_sub(); }
    close_sub();
  }
};

struct SubReader {
  FILE* f;
  uint64_t sub_left = 0;
  bool more = true, started = false, ok = true;
  void open_sub() {
    int32_t m = 0;
    ok = ok && fread(&m, 4, 1, f) == 1;
    more = m < 0;
    sub_left = (uint64_t)(m < 0 ? -(int64_t)m : (int64_t)m);
    started = true;
  }
  void close_sub() {
    int32_t m = 0;
    ok = ok && fread(&m, 4, 1, f) == 1;
  }
  void read(void* p, uint64_t n) {
    char* c = (char*)p;
    while (n > 0 && ok) {
      if (!started) open_sub();
      if (sub_left == 0) {
        if (!more) { ok = false; break; }       // record shorter than the I/O list
        close_sub();
        open_sub();
      }
      const uint64_t k = sub_left < n ? sub_left : n;
      if (c) { ok = ok && fread(c, 1, (size_t)k, f) == (size_t)k; c += k; }
      else ok = ok && fseek(f, (long)k, SEEK_CUR) == 0;
      n -= k; sub_left -= k;
    }
  }
  void finish() {               // skip what the I/O list did not consume
    if (!started) open_sub();
    for (;;) {
      if (sub_left) { ok = ok && fseek(f, (long)sub_left, SEEK_CUR) == 0; sub_left = 0; }
      close_sub();
      if (!more || !ok) break;
      open_sub();
    }
  }
};
}  // namespace

extern "C" {

void mrg_f77_set_max_subrecord(uint64_t bytes) { kMaxSub = bytes ? bytes : 2147483639ull; }

int mrg_f77_write_record(void* file, const void* const* parts, const uint64_t* bytes, int32_t nparts) {
  if (!file || nparts < 0) return MRG_ERR_ARG;
  SubWriter w{(FILE*)file, 0};
  for (int i = 0; i < nparts; i++) w.total += bytes[i];
  for (int i = 0; i < nparts; i++) w.write(parts[i], bytes[i]);
  w.finish();
  return w.ok ? MRG_OK : MRG_ERR_STATE;
}

int mrg_f77_read_record(void* file, void* const* parts, const uint64_t* bytes, int32_t nparts) {
  if (!file || nparts < 0) return MRG_ERR_ARG;
  SubReader r{(FILE*)file};
  for (int i = 0; i < nparts; i++) r.read(parts ? parts[i] : nullptr, bytes[i]);
  r.finish();
  return r.ok ? MRG_OK : MRG_ERR_STATE;
}

int mrg_restart_append_particles(mrg_ctx* ctx, const char* path, double qmulti, double wmulti, double qmulte, double wmulte,
                                 int64_t npr, int64_t np0, int64_t first, int64_t stride) {
  if (!ctx || !path || npr < 0 || np0 < npr || first < 1 || stride < 1) return MRG_ERR_ARG;
  FILE* f = fopen(path, "ab");
  if (!f) return MRG_ERR_STATE;
  int rc = MRG_OK;
  {
    const double q[4] = {qmulti, wmulti, qmulte, wmulte};
    const void* p[1] = {q};
    const uint64_t b[1] = {sizeof(q)};
    rc = mrg_f77_write_record(f, p, b, 1);                               // F:9722
  }
  if (!rc) {
    const int32_t n32 = (int32_t)npr;                                    // integer(C_INT) npr, F:9723
    const void* p[1] = {&n32};
    const uint64_t b[1] = {4};
    rc = mrg_f77_write_record(f, p, b, 1);
  }
  // one species = one record of six np0-long arrays: the library hands out the owned entries in l order, everything else
  // is zero (what the reference's zeroed scratch arrays hold for entries nobody owns, F:9622-9642)
  for (int ksp = 1; ksp <= 2 && !rc; ksp++) {
    std::vector<double> host[6];
    for (auto& h : host) h.assign((size_t)np0, 0.0);
    if (mrg_num_local(ctx, ksp) > 0)
      rc = mrg_download_particles(ctx, ksp, host[0].data(), host[1].data(), host[2].data(), host[3].data(), host[4].data(),
                                  host[5].data(), npr, first, stride);
    if (rc) break;
    const void* p[6];
    uint64_t b[6];
    for (int k = 0; k < 6; k++) { p[k] = host[k].data(); b[k] = (uint64_t)np0 * 8ull; }
    rc = mrg_f77_write_record(f, p, b, 6);                               // F:9724 / F:9725
  }
  if (fclose(f) != 0 && !rc) rc = MRG_ERR_STATE;
  return rc;
}

int mrg_restart_read_particles(mrg_ctx* ctx, const char* path, int32_t skip_records, double qw[4], int64_t* npr, int64_t np0,
                               int64_t first, int64_t stride) {
  if (!ctx || !path || !qw || !npr || np0 < 0 || first < 1 || stride < 1) return MRG_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return MRG_ERR_STATE;
  int rc = MRG_OK;
  for (int i = 0; i < skip_records && !rc; i++) rc = mrg_f77_read_record(f, nullptr, nullptr, 0);
  if (!rc) { void* p[1] = {qw}; const uint64_t b[1] = {32}; rc = mrg_f77_read_record(f, p, b, 1); }
  int32_t n32 = 0;
  if (!rc) { void* p[1] = {&n32}; const uint64_t b[1] = {4}; rc = mrg_f77_read_record(f, p, b, 1); }
  *npr = n32;
  if (!rc && (n32 < 0 || n32 > np0)) rc = MRG_ERR_STATE;
  for (int ksp = 1; ksp <= 2 && !rc; ksp++) {
    std::vector<double> host[6];
    for (auto& h : host) h.assign((size_t)np0, 0.0);
    void* p[6];
    uint64_t b[6];
    for (int k = 0; k < 6; k++) { p[k] = host[k].data(); b[k] = (uint64_t)np0 * 8ull; }
    rc = mrg_f77_read_record(f, p, b, 6);
    if (!rc) rc = mrg_upload_particles(ctx, ksp, host[0].data(), host[1].data(), host[2].data(), host[3].data(), host[4].data(),
                                       host[5].data(), n32, first, stride);
  }
  fclose(f);
  return rc;
}

}  // extern "C"
