// rtb200_jpeg.hpp — a small baseline-JPEG (SOF0/SOF1, Huffman, 8-bit) decoder, plus P6 PPM.
//
// Why it exists: the reference loads textures through stb_image (core/rtw_stb_image.hpp:79),
// which it does not vendor (_cmake/stb.cmake:6-9 fetches it from the network).  A drop-in
// must still open "earthmap.jpg", so the host side carries its own decoder.  The integer
// arithmetic follows the published IJG algorithms — the LL&M "islow" 8x8 inverse DCT with
// 13-bit constants, the 16-bit fixed-point YCbCr->RGB tables, triangle-filter ("fancy")
// chroma upsampling — so that the texels can be pinned bit-exactly against libjpeg-turbo
// (PIL) in tests/test_jpeg.py.  Parity against stb's own IDCT is unpinned.
#ifndef RTB200_JPEG_HPP
#define RTB200_JPEG_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace rtb200 {

class jpeg_decoder {
 public:
  bool decode(const uint8_t* data, size_t size, std::vector<uint8_t>& rgb, int& width, int& height) {
    d_ = data, n_ = size, pos_ = 0;
    if (n_ < 4 || d_[0] != 0xFF || d_[1] != 0xD8) return fail("not a JPEG");
    pos_ = 2;
    bool have_frame = false;
    while (pos_ + 4 <= n_) {
      if (d_[pos_] != 0xFF) return fail("marker expected");
      uint8_t m = d_[pos_ + 1];
      pos_ += 2;
      if (m == 0xFF) { pos_ -= 1; continue; }  // fill byte
      if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
      if (m == 0xD9) break;
      if (pos_ + 2 > n_) return fail("truncated");
      size_t len = (size_t(d_[pos_]) << 8) | d_[pos_ + 1];
      if (len < 2 || pos_ + len > n_) return fail("bad segment length");
      const uint8_t* seg = d_ + pos_ + 2;
      size_t seglen = len - 2;
      switch (m) {
        case 0xC0: case 0xC1:
          if (!read_sof(seg, seglen)) return false;
          have_frame = true;
          break;
        case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB:
        case 0xCD: case 0xCE: case 0xCF:
          return fail("only baseline / extended-sequential Huffman JPEG is supported");
        case 0xC4:
          if (!read_dht(seg, seglen)) return false;
          break;
        case 0xDB:
          if (!read_dqt(seg, seglen)) return false;
          break;
        case 0xDD:
          if (seglen < 2) return fail("bad DRI");
          restart_interval_ = (seg[0] << 8) | seg[1];
          break;
        case 0xDA:
          if (!have_frame) return fail("SOS before SOF");
          pos_ += len;
          if (!read_scan(seg, seglen)) return false;
          finish(rgb);
          width = width_, height = height_;
          return true;
        default: break;  // APPn, COM, ...
      }
      pos_ += len;
    }
    return fail("no scan found");
  }
  const char* error() const { return err_; }

 private:
  struct component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocks_w = 0, blocks_h = 0;  // padded to whole MCUs
    int dc_pred = 0;
    std::vector<uint8_t> plane;      // blocks_w*8 x blocks_h*8 samples
  };
  struct huff {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    uint16_t lookup[512];  // 9-bit fast path: (length << 8) | symbol, 0 = miss
  };

  bool fail(const char* why) { err_ = why; return false; }

  bool read_sof(const uint8_t* s, size_t n) {
    if (n < 6 || s[0] != 8) return fail("only 8-bit precision");
    height_ = (s[1] << 8) | s[2];
    width_ = (s[3] << 8) | s[4];
    int nc = s[5];
    if (width_ <= 0 || height_ <= 0) return fail("bad dimensions");
    if ((nc != 1 && nc != 3) || n < size_t(6 + 3 * nc)) return fail("only 1 or 3 components");
    comps_.assign(size_t(nc), component());
    hmax_ = vmax_ = 1;
    for (int i = 0; i < nc; i++) {
      component& c = comps_[size_t(i)];
      c.id = s[6 + 3 * i];
      c.h = s[7 + 3 * i] >> 4;
      c.v = s[7 + 3 * i] & 15;
      c.tq = s[8 + 3 * i];
      if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) return fail("unsupported sampling factors");
      if (c.h > hmax_) hmax_ = c.h;
      if (c.v > vmax_) vmax_ = c.v;
    }
    if (nc == 3 && (comps_[1].h != 1 || comps_[1].v != 1 || comps_[2].h != 1 || comps_[2].v != 1 ||
                    comps_[0].h != hmax_ || comps_[0].v != vmax_))
      return fail("unsupported chroma layout");
    if (nc == 1) comps_[0].h = comps_[0].v = hmax_ = vmax_ = 1;  // single component: MCU is one block
    mcus_x_ = (width_ + 8 * hmax_ - 1) / (8 * hmax_);
    mcus_y_ = (height_ + 8 * vmax_ - 1) / (8 * vmax_);
    for (auto& c : comps_) {
      c.blocks_w = mcus_x_ * c.h;
      c.blocks_h = mcus_y_ * c.v;
      c.plane.assign(size_t(c.blocks_w) * 8 * c.blocks_h * 8, 0);
    }
    return true;
  }

  bool read_dqt(const uint8_t* s, size_t n) {
    static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    size_t i = 0;
    while (i < n) {
      int pq = s[i] >> 4, tq = s[i] & 15;
      i++;
      if (tq > 3) return fail("bad DQT id");
      for (int k = 0; k < 64; k++) {
        if (i + (pq ? 2 : 1) > n) return fail("truncated DQT");
        int q = pq ? ((s[i] << 8) | s[i + 1]) : s[i];
        i += pq ? 2 : 1;
        quant_[tq][zz[k]] = q;  // store in natural order
      }
    }
    std::memcpy(zigzag_, zz, 64);
    return true;
  }

  bool read_dht(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
      if (i + 17 > n) return fail("truncated DHT");
      int tc = s[i] >> 4, th = s[i] & 15;
      if (tc > 1 || th > 3) return fail("bad DHT id");
      huff& h = tables_[tc][th];
      int total = 0;
      for (int l = 1; l <= 16; l++) { h.bits[l] = s[i + size_t(l)]; total += h.bits[l]; }
      i += 17;
      if (total > 256 || i + size_t(total) > n) return fail("bad DHT counts");
      std::memcpy(h.vals, s + i, size_t(total));
      i += size_t(total);
      // canonical code assignment (ITU T.81 Annex C / F.2.2.3)
      int code = 0, k = 0;
      std::memset(h.lookup, 0, sizeof h.lookup);
      for (int l = 1; l <= 16; l++) {
        h.valptr[l] = k;
        h.mincode[l] = code;
        // Kraft check: the canonical codes of length l must fit l bits (stb_image rejects such tables too); without it
        // an over-subscribed table (e.g. bits[1] = 200) walks `first` past the 512-entry lookup table
        if (code + h.bits[l] > (1 << l)) return fail("bad DHT: over-subscribed code lengths");
        for (int j = 0; j < h.bits[l]; j++, k++, code++) {
          if (l <= 9) {
            int first = code << (9 - l);
            for (int f = 0; f < (1 << (9 - l)); f++) h.lookup[first + f] = uint16_t((l << 8) | h.vals[k]);
          }
        }
        h.maxcode[l] = h.bits[l] ? code - 1 : -1;
        code <<= 1;
      }
      h.maxcode[17] = 0x7fffffff;
      h.present = true;
    }
    return true;
  }

  // ---- entropy-coded segment bit reader (0xFF00 unstuffing; stops at markers) --------
  void fill_bits() {
    while (bit_count_ <= 24) {
      int b = 0;
      if (!hit_marker_ && pos_ < n_) {
        b = d_[pos_];
        if (b == 0xFF) {
          int b2 = pos_ + 1 < n_ ? d_[pos_ + 1] : 0xD9;
          if (b2 == 0) pos_ += 2;
          else { hit_marker_ = true; b = 0; }
        } else {
          pos_++;
        }
      }
      bit_buf_ |= uint32_t(b) << (24 - bit_count_);
      bit_count_ += 8;
    }
  }
  int peek(int nbits) { if (bit_count_ < nbits) fill_bits(); return int(bit_buf_ >> (32 - nbits)); }
  void skip(int nbits) { bit_buf_ <<= nbits; bit_count_ -= nbits; }
  int receive_extend(int s) {
    if (!s) return 0;
    int v = peek(s);
    skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  int decode_symbol(const huff& h) {
    int look = peek(9);
    uint16_t e = h.lookup[look];
    if (e) { skip(e >> 8); return e & 255; }
    int code = peek(16);
    for (int l = 10; l <= 16; l++) {
      int c = code >> (16 - l);
      if (h.maxcode[l] >= 0 && c <= h.maxcode[l] && c >= h.mincode[l]) {
        const int at = h.valptr[l] + c - h.mincode[l];
        if (at < 0 || at > 255) break;  // (cannot happen for a table that passed read_dht; kept as a bound on vals[])
        skip(l);
        return h.vals[at];
      }
    }
    bad_code_ = true;
    skip(16);
    return 0;
  }

  bool read_scan(const uint8_t* s, size_t n) {
    if (n < 1) return fail("bad SOS");
    int ns = s[0];
    if (ns != int(comps_.size()) || n < size_t(1 + 2 * ns + 3)) return fail("only single-scan (interleaved) JPEG");
    for (int i = 0; i < ns; i++) {
      int cid = s[1 + 2 * i], t = s[2 + 2 * i];
      bool found = false;
      for (auto& c : comps_)
        if (c.id == cid) { c.td = t >> 4; c.ta = t & 15; found = true; }
      if (!found) return fail("SOS names an unknown component");
    }
    for (auto& c : comps_)
      if (c.td > 3 || c.ta > 3 || !tables_[0][c.td].present || !tables_[1][c.ta].present)
        return fail("missing Huffman table");
    bit_buf_ = 0, bit_count_ = 0, hit_marker_ = false, bad_code_ = false;
    int until_restart = restart_interval_;
    int coef[64];
    for (int my = 0; my < mcus_y_; my++) {
      for (int mx = 0; mx < mcus_x_; mx++) {
        if (restart_interval_ && until_restart == 0) {
          // byte-align, consume RSTn, reset predictors
          bit_buf_ = 0, bit_count_ = 0;
          while (pos_ + 1 < n_ && !(d_[pos_] == 0xFF && d_[pos_ + 1] != 0x00)) pos_++;  // skip pad bytes
          if (pos_ + 1 < n_ && d_[pos_ + 1] >= 0xD0 && d_[pos_ + 1] <= 0xD7) pos_ += 2;
          hit_marker_ = false;
          for (auto& c : comps_) c.dc_pred = 0;
          until_restart = restart_interval_;
        }
        for (auto& c : comps_) {
          for (int by = 0; by < c.v; by++)
            for (int bx = 0; bx < c.h; bx++) {
              decode_block(c, coef);
              int stride = c.blocks_w * 8;
              uint8_t* out = c.plane.data() + size_t((my * c.v + by) * 8) * stride + size_t((mx * c.h + bx) * 8);
              idct_islow(coef, quant_[c.tq], out, stride);
            }
        }
        if (restart_interval_) until_restart--;
      }
    }
    if (bad_code_) return fail("corrupt entropy-coded data");
    return true;
  }

  void decode_block(component& c, int* coef) {
    std::memset(coef, 0, 64 * sizeof(int));
    const huff& dc = tables_[0][c.td];
    const huff& ac = tables_[1][c.ta];
    int s = decode_symbol(dc);
    c.dc_pred += receive_extend(s & 15);
    coef[0] = c.dc_pred;
    for (int k = 1; k < 64;) {
      int rs = decode_symbol(ac);
      int r = rs >> 4, sz = rs & 15;
      if (sz == 0) {
        if (r != 15) break;  // EOB
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) { bad_code_ = true; break; }
      coef[zigzag_[k]] = receive_extend(sz);
      k++;
    }
  }

  // ---- IJG "islow" inverse DCT (Loeffler-Ligtenberg-Moschytz, CONST_BITS 13, PASS1_BITS 2)
  static int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
  static uint8_t clamp8(int v) { return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v)); }
  static void idct_1d(const int in[8], int out[8], int even_shift) {
    // even part
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * -15137;
    int tmp3 = z1 + z2 * 6270;
    z2 = in[0], z3 = in[4];
    int tmp0 = (z2 + z3) * (1 << even_shift);
    int tmp1 = (z2 - z3) * (1 << even_shift);
    int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    // odd part
    tmp0 = in[7], tmp1 = in[5], tmp2 = in[3], tmp3 = in[1];
    z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446, tmp1 *= 16819, tmp2 *= 25172, tmp3 *= 12299;
    z1 *= -7373, z2 *= -20995, z3 *= -16069, z4 *= -3196;
    z3 += z5, z4 += z5;
    tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
    out[0] = tmp10 + tmp3, out[7] = tmp10 - tmp3;
    out[1] = tmp11 + tmp2, out[6] = tmp11 - tmp2;
    out[2] = tmp12 + tmp1, out[5] = tmp12 - tmp1;
    out[3] = tmp13 + tmp0, out[4] = tmp13 - tmp0;
  }
  static void idct_islow(const int* coef, const int* q, uint8_t* out, int stride) {
    int ws[64];
    for (int col = 0; col < 8; col++) {  // pass 1: columns, scaled up by 2^PASS1_BITS
      int in[8], o[8];
      for (int r = 0; r < 8; r++) in[r] = coef[r * 8 + col] * q[r * 8 + col];
      idct_1d(in, o, 13);
      for (int r = 0; r < 8; r++) ws[r * 8 + col] = descale(o[r], 13 - 2);
    }
    for (int row = 0; row < 8; row++) {  // pass 2: rows, remove 2^(PASS1_BITS+3), level shift
      int o[8];
      idct_1d(ws + row * 8, o, 13);
      for (int c = 0; c < 8; c++) out[row * stride + c] = clamp8(descale(o[c], 13 + 2 + 3) + 128);
    }
  }

  // ---- upsampling + colour conversion ------------------------------------------------
  // Triangle-filter chroma upsampling as in the IJG decoder's "fancy" mode.
  static void upsample_h2(const uint8_t* in, int n_in, uint8_t* out) {
    if (n_in == 1) { out[0] = out[1] = in[0]; return; }
    out[0] = in[0];
    out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
    for (int i = 1; i < n_in - 1; i++) {
      int v = in[i] * 3;
      out[2 * i] = uint8_t((v + in[i - 1] + 1) >> 2);
      out[2 * i + 1] = uint8_t((v + in[i + 1] + 2) >> 2);
    }
    out[2 * n_in - 2] = uint8_t((in[n_in - 1] * 3 + in[n_in - 2] + 1) >> 2);
    out[2 * n_in - 1] = in[n_in - 1];
  }
  static void upsample_h2v2_row(const uint8_t* near_row, const uint8_t* far_row, int n_in, uint8_t* out) {
    // vertical 3:1 blend first (kept at x4 scale), then horizontal 3:1 with alternating bias
    std::vector<int> col(static_cast<size_t>(n_in), 0);
    for (int i = 0; i < n_in; i++) col[size_t(i)] = near_row[i] * 3 + far_row[i];
    if (n_in == 1) { out[0] = out[1] = uint8_t((col[0] * 4 + 8) >> 4); return; }
    out[0] = uint8_t((col[0] * 4 + 8) >> 4);
    out[1] = uint8_t((col[0] * 3 + col[1] + 7) >> 4);
    for (int i = 1; i < n_in - 1; i++) {
      out[2 * i] = uint8_t((col[size_t(i)] * 3 + col[size_t(i - 1)] + 8) >> 4);
      out[2 * i + 1] = uint8_t((col[size_t(i)] * 3 + col[size_t(i + 1)] + 7) >> 4);
    }
    out[2 * n_in - 2] = uint8_t((col[size_t(n_in - 1)] * 3 + col[size_t(n_in - 2)] + 8) >> 4);
    out[2 * n_in - 1] = uint8_t((col[size_t(n_in - 1)] * 4 + 7) >> 4);
  }

  void finish(std::vector<uint8_t>& rgb) {
    rgb.assign(size_t(width_) * height_ * 3, 0);
    if (comps_.size() == 1) {
      const component& y = comps_[0];
      int stride = y.blocks_w * 8;
      for (int j = 0; j < height_; j++)
        for (int i = 0; i < width_; i++) {
          uint8_t v = y.plane[size_t(j) * stride + i];
          uint8_t* p = &rgb[(size_t(j) * width_ + i) * 3];
          p[0] = p[1] = p[2] = v;
        }
      return;
    }
    // chroma planes at full resolution
    const component& Y = comps_[0];
    const int ys = Y.blocks_w * 8;
    std::vector<uint8_t> full[2];
    const uint8_t* cptr[2];
    int cstride[2];
    for (int k = 0; k < 2; k++) {
      const component& c = comps_[size_t(k + 1)];
      const int cs = c.blocks_w * 8;
      if (hmax_ == 1 && vmax_ == 1) { cptr[k] = c.plane.data(); cstride[k] = cs; continue; }
      // true (unpadded) chroma dimensions
      const int cw = (width_ + hmax_ - 1) / hmax_, ch = (height_ + vmax_ - 1) / vmax_;
      const int fw = cw * hmax_;
      full[k].assign(size_t(fw) * (size_t(ch) * vmax_), 0);
      for (int j = 0; j < ch * vmax_; j++) {
        uint8_t* out = &full[k][size_t(j) * fw];
        if (vmax_ == 1) {
          upsample_h2(&c.plane[size_t(j) * cs], cw, out);
        } else {
          const int src = j >> 1;
          const int other = (j & 1) ? (src + 1 < ch ? src + 1 : src) : (src > 0 ? src - 1 : src);
          if (hmax_ == 2) {
            upsample_h2v2_row(&c.plane[size_t(src) * cs], &c.plane[size_t(other) * cs], cw, out);
          } else {  // h1v2: vertical triangle only
            for (int i = 0; i < cw; i++)
              out[i] = uint8_t((c.plane[size_t(src) * cs + i] * 3 + c.plane[size_t(other) * cs + i] + ((j & 1) ? 2 : 1)) >> 2);
          }
        }
      }
      cptr[k] = full[k].data();
      cstride[k] = fw;
    }
    // IJG fixed-point YCbCr -> RGB (SCALEBITS 16)
    int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
    for (int i = 0; i < 256; i++) {
      int x = i - 128;
      cr_r[i] = (91881 * x + 32768) >> 16;
      cb_b[i] = (116130 * x + 32768) >> 16;
      cr_g[i] = -46802 * x;
      cb_g[i] = -22554 * x + 32768;
    }
    for (int j = 0; j < height_; j++)
      for (int i = 0; i < width_; i++) {
        int y = Y.plane[size_t(j) * ys + i];
        int cb = cptr[0][size_t(j) * cstride[0] + i], cr = cptr[1][size_t(j) * cstride[1] + i];
        uint8_t* p = &rgb[(size_t(j) * width_ + i) * 3];
        p[0] = clamp8(y + cr_r[cr]);
        p[1] = clamp8(y + ((cb_g[cb] + cr_g[cr]) >> 16));
        p[2] = clamp8(y + cb_b[cb]);
      }
  }

  const uint8_t* d_ = nullptr;
  size_t n_ = 0, pos_ = 0;
  const char* err_ = "";
  int width_ = 0, height_ = 0, hmax_ = 1, vmax_ = 1, mcus_x_ = 0, mcus_y_ = 0;
  int restart_interval_ = 0;
  std::vector<component> comps_;
  int quant_[4][64] = {{0}};
  uint8_t zigzag_[64] = {0};
  huff tables_[2][4];
  uint32_t bit_buf_ = 0;
  int bit_count_ = 0;
  bool hit_marker_ = false, bad_code_ = false;
};

inline bool read_file(const std::string& path, std::vector<uint8_t>& bytes) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n <= 0) { std::fclose(f); return false; }
  bytes.resize(size_t(n));
  size_t got = std::fread(bytes.data(), 1, size_t(n), f);
  std::fclose(f);
  return got == size_t(n);
}

// JPEG (baseline) or binary PPM (P6, maxval 255) -> tightly packed RGB8.
inline bool load_image_rgb8(const std::string& path, std::vector<uint8_t>& rgb, int& w, int& h) {
  std::vector<uint8_t> bytes;
  if (!read_file(path, bytes)) return false;
  if (bytes.size() > 2 && bytes[0] == 'P' && bytes[1] == '6') {
    size_t pos = 2;
    int vals[3], got = 0;
    while (got < 3 && pos < bytes.size()) {
      while (pos < bytes.size() && (bytes[pos] == ' ' || bytes[pos] == '\n' || bytes[pos] == '\r' || bytes[pos] == '\t')) pos++;
      if (pos < bytes.size() && bytes[pos] == '#') { while (pos < bytes.size() && bytes[pos] != '\n') pos++; continue; }
      int v = 0, digits = 0;
      while (pos < bytes.size() && bytes[pos] >= '0' && bytes[pos] <= '9') { v = v * 10 + (bytes[pos++] - '0'); digits++; }
      if (!digits) return false;
      vals[got++] = v;
    }
    pos++;  // single whitespace after maxval
    if (got < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0) return false;
    size_t need = size_t(vals[0]) * vals[1] * 3;
    if (pos + need > bytes.size()) return false;
    rgb.assign(bytes.begin() + long(pos), bytes.begin() + long(pos + need));
    w = vals[0], h = vals[1];
    return true;
  }
  jpeg_decoder dec;
  if (!dec.decode(bytes.data(), bytes.size(), rgb, w, h)) {
    std::fprintf(stderr, "rtb200: cannot decode '%s': %s\n", path.c_str(), dec.error());
    return false;
  }
  return true;
}

}  // namespace rtb200
#endif
