// AVIF (ISO-BMFF / HEIF / MIAF still image) muxer: AV1 OBU payload(s) -> .avif file bytes.
//
// Replaces the reference's call into the crates.io dependency avif-serialize 0.8.6 [EXT, not on
// disk]:  Aviffy::new().matrix_coefficients(..).premultiplied_alpha(..).set_exif(..)
//         .to_vec(&color, alpha, width, height, depth)
// (/root/reference/ravif/src/av1encoder.rs:457-473).  Colour primaries / transfer are not
// passed by ravif, so the muxer defaults apply: BT.709 primaries, sRGB transfer, full range -
// consistent with the ColorDescription ravif hands to the encoder (:407-411).
// Layout follows SURVEY.md Appendix D; accepted by libavif's strict parser (tests/test_avif.py).
#include "avif_mux.h"

#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace {
struct Buf {
  std::vector<uint8_t> d;
  void u8(uint32_t v) { d.push_back((uint8_t)v); }
  void u16(uint32_t v) { u8(v >> 8); u8(v); }
  void u32(uint32_t v) { u16(v >> 16); u16(v); }
  void tag(const char *t) { d.insert(d.end(), t, t + 4); }
  void bytes(const void *p, size_t n) { d.insert(d.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
  size_t begin(const char *type) { size_t at = d.size(); u32(0); tag(type); return at; }
  size_t begin_full(const char *type, int version, uint32_t flags) { size_t at = begin(type); u32(((uint32_t)version << 24) | flags); return at; }
  void end(size_t at) { uint32_t n = (uint32_t)(d.size() - at); d[at] = n >> 24; d[at + 1] = n >> 16; d[at + 2] = n >> 8; d[at + 3] = n; }
};

void av1c(Buf &b, int profile, int level, int high_bitdepth, int mono) {
  size_t at = b.begin("av1C");
  b.u8(0x81);
  b.u8((profile << 5) | (level & 31));
  const int ss = mono ? 1 : 0;   // 4:4:4 colour: 0,0 ; monochrome: 1,1
  b.u8((0 << 7) | (high_bitdepth << 6) | (0 << 5) | (mono << 4) | (ss << 3) | (ss << 2) | 0);
  b.u8(0);
  b.end(at);
}
}  // namespace

extern "C" int b200avif_mux(const B200AvifMuxInput *in, uint8_t **out, size_t *out_len) {
  if (!in || !in->color || !in->color_len || !out || !out_len) return 4;
  const bool has_alpha = in->alpha && in->alpha_len;
  const bool has_exif = in->exif && in->exif_len;
  const int n_items = 1 + (has_alpha ? 1 : 0) + (has_exif ? 1 : 0);
  const int exif_id = has_alpha ? 3 : 2;
  Buf b;
  size_t at = b.begin("ftyp");
  b.tag("avif"); b.u32(0); b.tag("avif"); b.tag("mif1"); b.tag("miaf");
  b.end(at);

  const size_t meta = b.begin_full("meta", 0, 0);
  at = b.begin_full("hdlr", 0, 0);
  b.u32(0); b.tag("pict"); b.u32(0); b.u32(0); b.u32(0); b.u8(0);
  b.end(at);
  at = b.begin_full("pitm", 0, 0);
  b.u16(1);
  b.end(at);
  // iloc: offsets patched once the mdat position is known
  at = b.begin_full("iloc", 0, 0);
  b.u8(0x44); b.u8(0x00);
  b.u16(n_items);
  size_t off_pos[3];
  const uint32_t lens[3] = {(uint32_t)in->color_len, (uint32_t)in->alpha_len, (uint32_t)(in->exif_len + 4)};
  int k = 0;
  for (int item = 1; item <= 3; item++) {
    if (item == 2 && !has_alpha) continue;
    if (item == 3 && !has_exif) continue;
    const int id = item == 3 ? exif_id : item;
    b.u16(id); b.u16(0); b.u16(1);
    off_pos[k] = b.d.size();
    b.u32(0); b.u32(lens[item - 1]);
    k++;
  }
  b.end(at);
  at = b.begin_full("iinf", 0, 0);
  b.u16(n_items);
  {
    size_t e = b.begin_full("infe", 2, 0);
    b.u16(1); b.u16(0); b.tag("av01"); b.bytes("Color", 6);
    b.end(e);
    if (has_alpha) {
      e = b.begin_full("infe", 2, 0);
      b.u16(2); b.u16(0); b.tag("av01"); b.bytes("Alpha", 6);
      b.end(e);
    }
    if (has_exif) {
      e = b.begin_full("infe", 2, 0);
      b.u16(exif_id); b.u16(0); b.tag("Exif"); b.bytes("Exif", 5);
      b.end(e);
    }
  }
  b.end(at);
  if (has_alpha || has_exif) {
    at = b.begin_full("iref", 0, 0);
    if (has_alpha) {
      size_t e = b.begin("auxl");
      b.u16(2); b.u16(1); b.u16(1);
      b.end(e);
      if (in->premultiplied_alpha) {
        e = b.begin("prem");
        b.u16(1); b.u16(1); b.u16(2);
        b.end(e);
      }
    }
    if (has_exif) {
      size_t e = b.begin("cdsc");
      b.u16(exif_id); b.u16(1); b.u16(1);
      b.end(e);
    }
    b.end(at);
  }
  at = b.begin("iprp");
  {
    const size_t ipco = b.begin("ipco");
    size_t e = b.begin_full("ispe", 0, 0);              // 1
    b.u32(in->width); b.u32(in->height);
    b.end(e);
    av1c(b, 1, in->color_level_idx, in->depth > 8, 0);  // 2
    e = b.begin_full("pixi", 0, 0);                     // 3
    b.u8(3); b.u8(in->depth); b.u8(in->depth); b.u8(in->depth);
    b.end(e);
    e = b.begin("colr");                                // 4
    b.tag("nclx"); b.u16(1); b.u16(13); b.u16(in->matrix_coefficients); b.u8(0x80);
    b.end(e);
    if (has_alpha) {
      av1c(b, 0, in->alpha_level_idx, in->depth > 8, 1);  // 5
      e = b.begin_full("pixi", 0, 0);                     // 6
      b.u8(1); b.u8(in->depth);
      b.end(e);
      e = b.begin_full("auxC", 0, 0);                     // 7
      static const char urn[] = "urn:mpeg:mpegB:cicp:systems:auxiliary:alpha";
      b.bytes(urn, sizeof(urn));
      b.end(e);
    }
    b.end(ipco);
    e = b.begin_full("ipma", 0, 0);
    b.u32(has_alpha ? 2 : 1);
    b.u16(1); b.u8(4); b.u8(1); b.u8(0x80 | 2); b.u8(3); b.u8(4);
    if (has_alpha) { b.u16(2); b.u8(4); b.u8(1); b.u8(0x80 | 5); b.u8(6); b.u8(7); }
    b.end(e);
  }
  b.end(at);
  b.end(meta);

  at = b.begin("mdat");
  const uint32_t o_color = (uint32_t)b.d.size();
  b.bytes(in->color, in->color_len);
  uint32_t o_alpha = 0, o_exif = 0;
  if (has_alpha) { o_alpha = (uint32_t)b.d.size(); b.bytes(in->alpha, in->alpha_len); }
  if (has_exif) { o_exif = (uint32_t)b.d.size(); b.u32(0); b.bytes(in->exif, in->exif_len); }
  b.end(at);
  const uint32_t offs[3] = {o_color, o_alpha, o_exif};
  k = 0;
  for (int item = 1; item <= 3; item++) {
    if (item == 2 && !has_alpha) continue;
    if (item == 3 && !has_exif) continue;
    const uint32_t o = offs[item - 1];
    uint8_t *p = &b.d[off_pos[k++]];
    p[0] = o >> 24; p[1] = o >> 16; p[2] = o >> 8; p[3] = o;
  }
  uint8_t *res = (uint8_t *)malloc(b.d.size());
  if (!res) return 3;
  memcpy(res, b.d.data(), b.d.size());
  *out = res;
  *out_len = b.d.size();
  return 0;
}
