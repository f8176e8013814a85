// weights_io.cpp — weight-file and image-file readers (host only, no CUDA).
//
// Replaces readSquezeNetKernel (MobileNet.c:31-47): the reference re-opens the text file and
// re-reads its first N tokens for every layer and truncates to int; here
// the file is parsed ONCE, sequentially, into the flat order the 29 layers consume
// (864, 288, 2048, ... 1024000 values; SURVEY App. A column nW) and kept as float.
//
// Two formats, auto-detected:
//   text   whitespace separated decimal tokens, exactly what fscanf("%s")+atof accepted
//          (MobileNet.c:41-42).  4 209 088 tokens = filters only (scale=1, shift=0);
//          4 209 088 + 2*(10 944+1000) tokens = filters, then scale[], then shift[].
//   binary "MNV1WTS1" | u64 n_weights | u64 n_channels | f32 weights[] | f32 scale[] | f32 shift[]
// and decode_image's replacement: a P6 PPM reader that skips the header (fixes D-15).
#include <cctype>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mnv1.h"

namespace mnv1 {

static const char kMagic[8] = {'M', 'N', 'V', '1', 'W', 'T', 'S', '1'};

static bool read_all(const char* path, std::vector<char>* out, std::string* err) {
  FILE* fp = fopen(path, "rb");
  if (!fp) { *err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (sz < 0) { fclose(fp); *err = "ftell failed"; return false; }
  out->resize((size_t)sz + 1);
  size_t got = fread(out->data(), 1, (size_t)sz, fp);
  fclose(fp);
  if (got != (size_t)sz) { *err = "short read"; return false; }
  (*out)[(size_t)sz] = '\0';
  return true;
}

// Parses the weight file into weights / scale / shift.  Returns MNV1_OK or MNV1_EIO.
int load_weight_file(const char* path, std::vector<float>* weights, std::vector<float>* scale,
                     std::vector<float>* shift, std::string* err) {
  const size_t NW = (size_t)MNV1_TOTAL_WEIGHTS, NC = (size_t)MNV1_BN_CHANNELS + MNV1_NUM_CLASSES;
  std::vector<char> raw;
  if (!read_all(path, &raw, err)) return MNV1_EIO;
  const size_t sz = raw.size() - 1;
  weights->assign(NW, 0.f);
  scale->assign(NC, 1.f);
  shift->assign(NC, 0.f);
  if (sz >= 24 && memcmp(raw.data(), kMagic, 8) == 0) {
    uint64_t nw, nc;
    memcpy(&nw, raw.data() + 8, 8);
    memcpy(&nc, raw.data() + 16, 8);
    if (nw != NW || (nc != 0 && nc != NC) || sz != 24 + 4 * (nw + 2 * nc)) {
      *err = "binary weight file has unexpected sizes";
      return MNV1_EIO;
    }
    memcpy(weights->data(), raw.data() + 24, 4 * NW);
    if (nc) {
      memcpy(scale->data(), raw.data() + 24 + 4 * NW, 4 * NC);
      memcpy(shift->data(), raw.data() + 24 + 4 * NW + 4 * NC, 4 * NC);
    }
    return MNV1_OK;
  }
  // text: strtof over the buffer (what fscanf("%s") + atof did, one pass instead of 29)
  std::vector<float> tok;
  tok.reserve(NW + 2 * NC);
  const char* p = raw.data();
  const char* end = raw.data() + sz;
  while (p < end) {
    while (p < end && isspace((unsigned char)*p)) ++p;
    if (p >= end) break;
    char* q = nullptr;
    float v = strtof(p, &q);
    if (q == p) {  // atof() of a non-number is 0: keep the reference's tolerance, skip the token
      while (p < end && !isspace((unsigned char)*p)) ++p;
      v = 0.f;
    } else {
      p = q;
    }
    tok.push_back(v);
    if (tok.size() > NW + 2 * NC) break;
  }
  if (tok.size() != NW && tok.size() != NW + 2 * NC) {
    char b[160];
    snprintf(b, sizeof b, "text weight file has %zu tokens, expected %zu or %zu", tok.size(), NW, NW + 2 * NC);
    *err = b;
    return MNV1_EIO;
  }
  memcpy(weights->data(), tok.data(), 4 * NW);
  if (tok.size() > NW) {
    memcpy(scale->data(), tok.data() + NW, 4 * NC);
    memcpy(shift->data(), tok.data() + NW + NC, 4 * NC);
  }
  return MNV1_OK;
}

int save_weight_file_bin(const char* path, const float* weights, const float* scale, const float* shift,
                         std::string* err) {
  const uint64_t NW = (uint64_t)MNV1_TOTAL_WEIGHTS, NC = (uint64_t)MNV1_BN_CHANNELS + MNV1_NUM_CLASSES;
  FILE* fp = fopen(path, "wb");
  if (!fp) { *err = std::string("cannot create ") + path + ": " + strerror(errno); return MNV1_EIO; }
  const bool has_bn = scale && shift;
  const uint64_t nc = has_bn ? NC : 0;
  bool ok = fwrite(kMagic, 1, 8, fp) == 8 && fwrite(&NW, 8, 1, fp) == 1 && fwrite(&nc, 8, 1, fp) == 1 &&
            fwrite(weights, 4, NW, fp) == NW;
  if (ok && has_bn) ok = fwrite(scale, 4, NC, fp) == NC && fwrite(shift, 4, NC, fp) == NC;
  fclose(fp);
  if (!ok) { *err = "short write"; return MNV1_EIO; }
  return MNV1_OK;
}

// decode_image (MobileNet.c:49-57) freads 150 528 bytes from offset 0, header included
// (App. C D-15).  This reader parses "P6 <w> <h> <maxval>\n" (comments allowed) first.
int read_ppm(const char* path, uint8_t* out, int height, int width, std::string* err) {
  std::vector<char> raw;
  if (!read_all(path, &raw, err)) return MNV1_EIO;
  const size_t sz = raw.size() - 1;
  size_t pos = 0;
  auto next_token = [&](std::string* t) {
    t->clear();
    while (pos < sz) {
      if (raw[pos] == '#') { while (pos < sz && raw[pos] != '\n') ++pos; }
      else if (isspace((unsigned char)raw[pos])) ++pos;
      else break;
    }
    while (pos < sz && !isspace((unsigned char)raw[pos])) t->push_back(raw[pos++]);
    return !t->empty();
  };
  std::string magic, sw, sh, smax;
  if (!next_token(&magic) || magic != "P6" || !next_token(&sw) || !next_token(&sh) || !next_token(&smax)) {
    *err = "not a binary PPM (P6)";
    return MNV1_EIO;
  }
  ++pos;  // single whitespace after maxval
  const int w = atoi(sw.c_str()), h = atoi(sh.c_str()), mx = atoi(smax.c_str());
  if (w != width || h != height || mx != 255) {
    char b[128];
    snprintf(b, sizeof b, "PPM is %dx%d max %d, expected %dx%d max 255", w, h, mx, width, height);
    *err = b;
    return MNV1_EIO;
  }
  const size_t need = (size_t)w * h * 3;
  if (sz - pos < need) { *err = "PPM payload truncated"; return MNV1_EIO; }
  memcpy(out, raw.data() + pos, need);
  return MNV1_OK;
}

}  // namespace mnv1
