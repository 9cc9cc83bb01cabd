// apd_jpeg.cpp -- JPEG decoder for the drop-in's ReadImage / ReadImageColor (APD.cpp:137-160, 1092: cv::imread).
//
// The reference reads its inputs with cv::imread, i.e. libjpeg(-turbo) with default settings.  To feed the kernels the
// SAME pixels, this decoder restates the arithmetic those defaults select (published algorithms of the IJG library):
//   * entropy decoding: baseline / extended sequential Huffman (SOF0, SOF1) and progressive Huffman (SOF2), 8 bit,
//     restart intervals, interleaved and non-interleaved scans;
//   * inverse DCT: the "islow" integer IDCT (Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 extra bits in pass 1);
//   * IMREAD_GRAYSCALE of a colour JPEG = the luma plane itself (out_color_space = JCS_GRAYSCALE);
//   * colour: "fancy" (triangle filter) chroma up-sampling for h2v1 / h2v2, box replication otherwise, and the 16-bit
//     fixed-point YCbCr -> RGB tables;
//   * EXIF orientation (cv::imread applies it unless IMREAD_IGNORE_ORIENTATION is passed; the reference does not pass it).
// Pinned against cv2 (libjpeg-turbo) in tests/test_host_io.py: grey and colour, 4:4:4 / 4:2:2 / 4:2:0, baseline,
// progressive, restart markers, odd sizes.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace apd {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[18], maxcode[18], valptr[18];
    int16_t fast[512];  // 9-bit look-ahead: (length << 8) | symbol, or -1
    void build() {
        int code = 0, k = 0;
        for (int i = 0; i < 512; ++i) fast[i] = -1;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                if (l <= 9) {
                    const int first = code << (9 - l);
                    for (int j = 0; j < (1 << (9 - l)); ++j) fast[first + j] = (int16_t)((l << 8) | vals[k]);
                }
            }
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;          // blocks per row / column of the padded (MCU-aligned) coefficient array
    int cw = 0, ch = 0;          // blocks actually covering the component (non-interleaved scans walk these)
    std::vector<int16_t> coef;   // [bh][bw][64], natural order
    std::vector<uint8_t> plane;  // [bh*8][bw*8] samples after the IDCT
};

struct Decoder {
    const uint8_t *d;
    size_t n, pos = 0;
    int W = 0, H = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    bool progressive = false, got_sof = false;
    Component comp[4];
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    int restart = 0, orientation = 1;
    bool adobe = false;
    int adobe_transform = -1;
    // bit reader
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;
    int eobrun = 0;

    Decoder(const uint8_t *data, size_t size) : d(data), n(size) { memset(qt, 0, sizeof(qt)); }

    int u8() { return pos < n ? d[pos++] : 0; }
    int u16() { const int a = u8(); return (a << 8) | u8(); }

    // ---- entropy-coded segment bits
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hit_marker && pos < n) {
                b = d[pos];
                if (b == 0xFF) {
                    const int b2 = pos + 1 < n ? d[pos + 1] : 0xD9;
                    if (b2 == 0) pos += 2;
                    else { hit_marker = true; b = 0; }  // leave the marker in place, feed zeros
                } else {
                    pos++;
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int getbits(int nb) {
        if (nb == 0) return 0;
        if (bitcnt < nb) fill();
        const int v = (int)(bitbuf >> (32 - nb));
        bitbuf <<= nb;
        bitcnt -= nb;
        return v;
    }
    int getbit() { return getbits(1); }
    static int extend(int v, int nb) { return v < (1 << (nb - 1)) ? v - (1 << nb) + 1 : v; }
    int receive_extend(int nb) { return nb ? extend(getbits(nb), nb) : 0; }
    int decode(const Huff &h) {
        if (bitcnt < 16) fill();
        const int look = (int)(bitbuf >> 23);
        const int f = h.fast[look];
        if (f >= 0) {
            const int l = f >> 8;
            bitbuf <<= l;
            bitcnt -= l;
            return f & 255;
        }
        int code = (int)(bitbuf >> 22), l = 10;  // first 10 bits
        while (l <= 16 && code > h.maxcode[l]) {
            l++;
            code = (int)(bitbuf >> (32 - l));
        }
        if (l > 16) { bitbuf <<= 16; bitcnt -= 16; return 0; }
        bitbuf <<= l;
        bitcnt -= l;
        return h.vals[(h.valptr[l] + code - h.mincode[l]) & 255];
    }
    void reset_bits() { bitbuf = 0; bitcnt = 0; hit_marker = false; eobrun = 0; }

    // ---- markers
    bool parse_dqt(int len) {
        const size_t end = pos + len;
        while (pos < end) {
            const int pq = u8();
            const int t = pq & 15, prec = pq >> 4;
            if (t > 3) return false;
            for (int i = 0; i < 64; ++i) qt[t][kZigzag[i]] = (uint16_t)(prec ? u16() : u8());
            qt_present[t] = true;
        }
        return pos == end;
    }
    bool parse_dht(int len) {
        const size_t end = pos + len;
        while (pos < end) {
            const int tc = u8();
            const int t = tc & 15, cls = tc >> 4;
            if (t > 3 || cls > 1) return false;
            Huff &h = cls ? ac[t] : dc[t];
            int total = 0;
            h.bits[0] = 0;
            for (int i = 1; i <= 16; ++i) { h.bits[i] = (uint8_t)u8(); total += h.bits[i]; }
            if (total > 256) return false;
            for (int i = 0; i < total; ++i) h.vals[i] = (uint8_t)u8();
            h.present = true;
            h.build();
        }
        return pos == end;
    }
    bool parse_sof(int len, bool prog) {
        const size_t end = pos + len;
        if (u8() != 8) return false;  // 8-bit samples only
        H = u16(); W = u16(); ncomp = u8();
        if (W <= 0 || H <= 0 || (ncomp != 1 && ncomp != 3)) return false;
        for (int i = 0; i < ncomp; ++i) {
            comp[i].id = u8();
            const int hv = u8();
            comp[i].h = hv >> 4; comp[i].v = hv & 15; comp[i].tq = u8() & 3;
            if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4) return false;
            hmax = comp[i].h > hmax ? comp[i].h : hmax;
            vmax = comp[i].v > vmax ? comp[i].v : vmax;
        }
        mcux = (W + 8 * hmax - 1) / (8 * hmax);
        mcuy = (H + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < ncomp; ++i) {
            Component &c = comp[i];
            c.bw = mcux * c.h; c.bh = mcuy * c.v;
            const int sw = (W * c.h + hmax - 1) / hmax, sh = (H * c.v + vmax - 1) / vmax;
            c.cw = (sw + 7) / 8; c.ch = (sh + 7) / 8;
            c.coef.assign((size_t)c.bw * c.bh * 64, 0);
        }
        progressive = prog;
        got_sof = true;
        return pos == end;
    }
    void parse_app1(int len) {  // EXIF orientation tag 0x0112 of IFD0
        const size_t start = pos, end = pos + len;
        if (len >= 14 && !memcmp(d + pos, "Exif\0\0", 6)) {
            const uint8_t *t = d + pos + 6;
            const size_t tl = len - 6;
            const bool le = t[0] == 'I';
            auto r16 = [&](size_t o) -> int { return o + 2 <= tl ? (le ? t[o] | (t[o + 1] << 8) : (t[o] << 8) | t[o + 1]) : 0; };
            auto r32 = [&](size_t o) -> uint32_t {
                return o + 4 <= tl ? (le ? (uint32_t)t[o] | (uint32_t)t[o + 1] << 8 | (uint32_t)t[o + 2] << 16 | (uint32_t)t[o + 3] << 24
                                         : (uint32_t)t[o] << 24 | (uint32_t)t[o + 1] << 16 | (uint32_t)t[o + 2] << 8 | t[o + 3]) : 0;
            };
            if (r16(2) == 42) {
                const size_t ifd = r32(4);
                const int cnt = r16(ifd);
                for (int i = 0; i < cnt; ++i) {
                    const size_t e = ifd + 2 + (size_t)12 * i;
                    if (r16(e) == 0x0112) { const int o = r16(e + 8); if (o >= 1 && o <= 8) orientation = o; }
                }
            }
        }
        (void)start;
        pos = end;
    }

    // ---- block decoders (ITU T.81 F.2.2 / G.1.2, as in the IJG library)
    bool decode_block_baseline(int16_t *blk, Component &c, int &pred) {
        const int s = decode(dc[c.td]);
        pred += receive_extend(s);
        blk[0] = (int16_t)pred;
        const Huff &h = ac[c.ta];
        for (int k = 1; k < 64;) {
            const int rs = decode(h);
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r != 15) break;
                k += 16;
            } else {
                k += r;
                if (k > 63) return false;
                blk[kZigzag[k]] = (int16_t)receive_extend(sz);
                k++;
            }
        }
        return true;
    }
    void decode_dc_first(int16_t *blk, Component &c, int &pred, int al) {
        const int s = decode(dc[c.td]);
        pred += receive_extend(s);
        blk[0] = (int16_t)(pred * (1 << al));
    }
    void decode_dc_refine(int16_t *blk, int al) {
        if (getbit()) blk[0] |= (int16_t)(1 << al);
    }
    bool decode_ac_first(int16_t *blk, Component &c, int ss, int se, int al) {
        if (eobrun > 0) { eobrun--; return true; }
        const Huff &h = ac[c.ta];
        for (int k = ss; k <= se;) {
            const int rs = decode(h);
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r) eobrun += getbits(r);
                    break;
                }
                k += 16;
            } else {
                k += r;
                if (k > 63) return false;
                blk[kZigzag[k]] = (int16_t)(receive_extend(sz) * (1 << al));
                k++;
            }
        }
        return true;
    }
    bool decode_ac_refine(int16_t *blk, Component &c, int ss, int se, int al) {
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        const Huff &h = ac[c.ta];
        if (eobrun == 0) {
            for (; k <= se; k++) {
                const int rs = decode(h);
                int r = rs >> 4;
                const int sz = rs & 15;
                int val = 0;
                if (sz) {
                    val = getbit() ? p1 : m1;  // sz must be 1
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += getbits(r);
                    break;
                }
                // skip r zero-history coefficients, refining the non-zero ones on the way
                while (k <= se) {
                    int16_t *cf = &blk[kZigzag[k]];
                    if (*cf != 0) {
                        if (getbit() && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
                    } else {
                        if (--r < 0) break;
                    }
                    k++;
                }
                if (val && k <= se) blk[kZigzag[k]] = (int16_t)val;
            }
        }
        if (eobrun > 0) {
            for (; k <= se; k++) {
                int16_t *cf = &blk[kZigzag[k]];
                if (*cf != 0 && getbit() && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
            }
            eobrun--;
        }
        return true;
    }

    bool process_restart(int &count, int preds[4]) {
        if (restart == 0) return true;
        if (--count > 0) return true;
        // expect an RSTn marker
        reset_bits();
        while (pos + 1 < n && !(d[pos] == 0xFF && d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7)) {
            if (d[pos] == 0xFF && d[pos + 1] != 0 && d[pos + 1] != 0xFF) return true;  // some other marker: let the caller see it
            pos++;
        }
        if (pos + 1 < n) pos += 2;
        count = restart;
        preds[0] = preds[1] = preds[2] = preds[3] = 0;
        return true;
    }

    bool parse_sos(int len) {
        const size_t end = pos + len;
        const int ns = u8();
        if (ns < 1 || ns > ncomp) return false;
        int idx[4];
        for (int i = 0; i < ns; ++i) {
            const int cid = u8(), t = u8();
            idx[i] = -1;
            for (int j = 0; j < ncomp; ++j) if (comp[j].id == cid) idx[i] = j;
            if (idx[i] < 0) return false;
            comp[idx[i]].td = t >> 4; comp[idx[i]].ta = t & 15;
            if (comp[idx[i]].td > 3 || comp[idx[i]].ta > 3) return false;
        }
        const int ss = u8(), se = u8(), ahl = u8();
        const int ah = ahl >> 4, al = ahl & 15;
        if (pos != end) return false;
        if (!progressive && (ss != 0 || se != 63)) {
            if (!(ss == 0 && se == 63)) return false;
        }
        reset_bits();
        int preds[4] = {0, 0, 0, 0};
        int count = restart;
        auto one_block = [&](Component &c, int ci, int bx, int by) -> bool {
            int16_t *blk = &c.coef[((size_t)by * c.bw + bx) * 64];
            if (!progressive) return decode_block_baseline(blk, c, preds[ci]);
            if (ss == 0) {
                if (ah == 0) decode_dc_first(blk, c, preds[ci], al);
                else decode_dc_refine(blk, al);
                return true;
            }
            return ah == 0 ? decode_ac_first(blk, c, ss, se, al) : decode_ac_refine(blk, c, ss, se, al);
        };
        if (ns == 1) {  // non-interleaved: the component's own block raster
            Component &c = comp[idx[0]];
            for (int by = 0; by < c.ch; ++by)
                for (int bx = 0; bx < c.cw; ++bx) {
                    if (!one_block(c, idx[0], bx, by)) return false;
                    process_restart(count, preds);
                }
        } else {
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    for (int i = 0; i < ns; ++i) {
                        Component &c = comp[idx[i]];
                        for (int y = 0; y < c.v; ++y)
                            for (int x = 0; x < c.h; ++x)
                                if (!one_block(c, idx[i], mx * c.h + x, my * c.v + y)) return false;
                    }
                    process_restart(count, preds);
                }
        }
        // drop what is left of the entropy-coded segment (padding bits)
        reset_bits();
        return true;
    }

    // ---- islow inverse DCT
    static inline uint8_t range_limit(int v) {  // sample_range_limit + CENTERJSAMPLE indexed by (v & RANGE_MASK)
        const int i = v & 1023;
        if (i < 128) return (uint8_t)(i + 128);
        if (i < 512) return 255;
        if (i < 896) return 0;
        return (uint8_t)(i - 896);
    }
    static void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, int stride) {
        const int CB = 13, P1 = 2;
        const int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137,
                  F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
        int ws[64];
        auto descale = [](long x, int nb) -> int { return (int)((x + (1L << (nb - 1))) >> nb); };
        for (int c = 0; c < 8; ++c) {
            auto D = [&](int r) -> long { return (long)in[r * 8 + c] * q[r * 8 + c]; };
            if (in[8 + c] == 0 && in[16 + c] == 0 && in[24 + c] == 0 && in[32 + c] == 0 && in[40 + c] == 0 && in[48 + c] == 0 &&
                in[56 + c] == 0) {
                const int dcv = (int)(D(0) * (1 << P1));
                for (int r = 0; r < 8; ++r) ws[r * 8 + c] = dcv;
                continue;
            }
            long z2 = D(2), z3 = D(6);
            long z1 = (z2 + z3) * F0541;
            long tmp2 = z1 + z3 * (-F1847);
            long tmp3 = z1 + z2 * F0765;
            z2 = D(0); z3 = D(4);
            long tmp0 = (z2 + z3) * (1L << CB);
            long tmp1 = (z2 - z3) * (1L << CB);
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = D(7); tmp1 = D(5); tmp2 = D(3); tmp3 = D(1);
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            ws[0 * 8 + c] = descale(tmp10 + tmp3, CB - P1);
            ws[7 * 8 + c] = descale(tmp10 - tmp3, CB - P1);
            ws[1 * 8 + c] = descale(tmp11 + tmp2, CB - P1);
            ws[6 * 8 + c] = descale(tmp11 - tmp2, CB - P1);
            ws[2 * 8 + c] = descale(tmp12 + tmp1, CB - P1);
            ws[5 * 8 + c] = descale(tmp12 - tmp1, CB - P1);
            ws[3 * 8 + c] = descale(tmp13 + tmp0, CB - P1);
            ws[4 * 8 + c] = descale(tmp13 - tmp0, CB - P1);
        }
        for (int r = 0; r < 8; ++r) {
            const int *w = &ws[r * 8];
            uint8_t *o = out + (size_t)r * stride;
            long z2 = w[2], z3 = w[6];
            long z1 = (z2 + z3) * F0541;
            long tmp2 = z1 + z3 * (-F1847);
            long tmp3 = z1 + z2 * F0765;
            long tmp0 = ((long)w[0] + w[4]) * (1L << CB);
            long tmp1 = ((long)w[0] - w[4]) * (1L << CB);
            const long tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
            z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
            long z4 = tmp1 + tmp3;
            const long z5 = (z3 + z4) * F1175;
            tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
            z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390;
            z3 += z5; z4 += z5;
            tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
            const int S = CB + P1 + 3;
            o[0] = range_limit(descale(tmp10 + tmp3, S));
            o[7] = range_limit(descale(tmp10 - tmp3, S));
            o[1] = range_limit(descale(tmp11 + tmp2, S));
            o[6] = range_limit(descale(tmp11 - tmp2, S));
            o[2] = range_limit(descale(tmp12 + tmp1, S));
            o[5] = range_limit(descale(tmp12 - tmp1, S));
            o[3] = range_limit(descale(tmp13 + tmp0, S));
            o[4] = range_limit(descale(tmp13 - tmp0, S));
        }
    }
    void idct_component(Component &c) {
        const int pw = c.bw * 8;
        c.plane.assign((size_t)pw * c.bh * 8, 0);
        for (int by = 0; by < c.bh; ++by)
            for (int bx = 0; bx < c.bw; ++bx)
                idct_islow(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &c.plane[(size_t)by * 8 * pw + bx * 8], pw);
    }

    bool parse() {
        if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
        pos = 2;
        while (pos + 4 <= n) {
            if (d[pos] != 0xFF) { pos++; continue; }
            const int m = d[pos + 1];
            if (m == 0xFF) { pos++; continue; }
            pos += 2;
            if (m == 0xD9) break;
            if (m == 0x00 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            const int len = u16() - 2;
            if (len < 0 || pos + len > n) return false;
            switch (m) {
                case 0xDB: if (!parse_dqt(len)) return false; break;
                case 0xC4: if (!parse_dht(len)) return false; break;
                case 0xC0: case 0xC1: if (!parse_sof(len, false)) return false; break;
                case 0xC2: if (!parse_sof(len, true)) return false; break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return false;  // lossless / hierarchical / arithmetic coding
                case 0xDD: restart = u16(); break;
                case 0xE1: parse_app1(len); break;
                case 0xEE:
                    if (len >= 12 && !memcmp(d + pos, "Adobe", 5)) { adobe = true; adobe_transform = d[pos + 11]; }
                    pos += len;
                    break;
                case 0xDA:
                    if (!got_sof || !parse_sos(len)) return false;
                    break;
                default: pos += len; break;
            }
        }
        return got_sof;
    }
};

// h2v1 "fancy" up-sampling of one row (jdsample.c h2v1_fancy_upsample)
void upsample_h2v1_row(const uint8_t *in, int inw, uint8_t *out) {
    if (inw == 1) { out[0] = out[1] = in[0]; return; }
    int v = in[0];
    out[0] = (uint8_t)v;
    out[1] = (uint8_t)((v * 3 + in[1] + 2) >> 2);
    for (int i = 1; i < inw - 1; ++i) {
        v = in[i] * 3;
        out[2 * i] = (uint8_t)((v + in[i - 1] + 1) >> 2);
        out[2 * i + 1] = (uint8_t)((v + in[i + 1] + 2) >> 2);
    }
    v = in[inw - 1];
    out[2 * (inw - 1)] = (uint8_t)((v * 3 + in[inw - 2] + 1) >> 2);
    out[2 * (inw - 1) + 1] = (uint8_t)v;
}
// h2v2 "fancy" up-sampling: one output row from the nearer (in0, weight 3) and the farther (in1, weight 1) input row
void upsample_h2v2_row(const uint8_t *in0, const uint8_t *in1, int inw, uint8_t *out) {
    if (inw == 1) { const int s = in0[0] * 3 + in1[0]; out[0] = (uint8_t)((s * 4 + 8) >> 4); out[1] = (uint8_t)((s * 4 + 7) >> 4); return; }
    int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
    out[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
    out[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
    lastcol = thiscol; thiscol = nextcol;
    for (int i = 1; i < inw - 1; ++i) {
        nextcol = in0[i + 1] * 3 + in1[i + 1];
        out[2 * i] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
        out[2 * i + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
        lastcol = thiscol; thiscol = nextcol;
    }
    out[2 * (inw - 1)] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
    out[2 * (inw - 1) + 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
}

// full-resolution plane of a chroma component (libjpeg default: fancy up-sampling for h2v1 and h2v2, replication otherwise)
void upsample_component(const Decoder &J, const Component &c, std::vector<uint8_t> &full, int &fw) {
    const int hs = J.hmax / c.h, vs = J.vmax / c.v;
    const int pw = c.bw * 8;
    const int sw = (J.W * c.h + J.hmax - 1) / J.hmax, sh = (J.H * c.v + J.vmax - 1) / J.vmax;  // down-sampled size
    fw = sw * hs;
    const int fh = sh * vs;
    full.assign((size_t)fw * fh, 0);
    const bool exact = (J.hmax % c.h == 0) && (J.vmax % c.v == 0);
    if (exact && hs == 1 && vs == 1) {
        for (int y = 0; y < sh; ++y) memcpy(&full[(size_t)y * fw], &c.plane[(size_t)y * pw], sw);
    } else if (exact && hs == 2 && vs == 1 && sw > 2) {
        for (int y = 0; y < sh; ++y) upsample_h2v1_row(&c.plane[(size_t)y * pw], sw, &full[(size_t)y * fw]);
    } else if (exact && hs == 2 && vs == 2 && sw > 2) {
        // libjpeg's main controller supplies the context rows; at the image top / bottom the missing neighbour is the first /
        // last REAL sample row duplicated (jdmainct.c make_funny_pointers / set_bottom_pointers), not the block padding
        for (int y = 0; y < sh; ++y) {
            const uint8_t *cur = &c.plane[(size_t)y * pw];
            const uint8_t *up = &c.plane[(size_t)(y > 0 ? y - 1 : 0) * pw];
            const uint8_t *dn = &c.plane[(size_t)(y + 1 < sh ? y + 1 : sh - 1) * pw];
            upsample_h2v2_row(cur, up, sw, &full[(size_t)(2 * y) * fw]);
            upsample_h2v2_row(cur, dn, sw, &full[(size_t)(2 * y + 1) * fw]);
        }
    } else if (exact && hs == 1 && vs == 2) {  // libjpeg-turbo h1v2_fancy_upsample (4:4:0)
        for (int y = 0; y < sh; ++y) {
            const uint8_t *cur = &c.plane[(size_t)y * pw];
            const uint8_t *up = &c.plane[(size_t)(y > 0 ? y - 1 : 0) * pw];
            const uint8_t *dn = &c.plane[(size_t)(y + 1 < sh ? y + 1 : sh - 1) * pw];
            for (int x = 0; x < sw; ++x) {
                full[(size_t)(2 * y) * fw + x] = (uint8_t)((cur[x] * 3 + up[x] + 1) >> 2);
                full[(size_t)(2 * y + 1) * fw + x] = (uint8_t)((cur[x] * 3 + dn[x] + 2) >> 2);
            }
        }
    } else {  // integral replication (h4v*, small widths, ...)
        for (int y = 0; y < fh; ++y)
            for (int x = 0; x < fw; ++x) full[(size_t)y * fw + x] = c.plane[(size_t)(y / vs) * pw + x / hs];
    }
}

// cv::imread's EXIF handling: rotate / flip so that the image is upright
void apply_orientation(int o, int &w, int &h, int ch, std::vector<uint8_t> &pix) {
    if (o <= 1 || o > 8) return;
    const int W = w, H = h;
    const bool swap = o >= 5;
    const int ow = swap ? H : W, oh = swap ? W : H;
    std::vector<uint8_t> out((size_t)ow * oh * ch);
    for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x) {
            int sx = x, sy = y;
            switch (o) {
                case 2: sx = W - 1 - x; sy = y; break;             // mirror horizontal
                case 3: sx = W - 1 - x; sy = H - 1 - y; break;     // rotate 180
                case 4: sx = x; sy = H - 1 - y; break;             // mirror vertical
                case 5: sx = y; sy = x; break;                     // transpose
                case 6: sx = y; sy = H - 1 - x; break;             // rotate 90 CW
                case 7: sx = W - 1 - y; sy = H - 1 - x; break;     // transverse
                case 8: sx = W - 1 - y; sy = x; break;             // rotate 270 CW
            }
            memcpy(&out[((size_t)y * ow + x) * ch], &pix[((size_t)sy * W + sx) * ch], ch);
        }
    pix.swap(out);
    w = ow; h = oh;
}

}  // namespace

// want_color = false: one channel (luma, as cv::imread(IMREAD_GRAYSCALE)); true: interleaved RGB
bool decode_jpeg(const std::vector<uint8_t> &f, bool want_color, int &w, int &h, int &channels, std::vector<uint8_t> &pix) {
    if (f.size() < 4 || f[0] != 0xFF || f[1] != 0xD8) return false;
    Decoder J(f.data(), f.size());
    if (!J.parse()) return false;
    for (int i = 0; i < J.ncomp; ++i) if (!J.qt_present[J.comp[i].tq]) return false;
    w = J.W; h = J.H;
    const bool ycc = J.ncomp == 3 && !(J.adobe && J.adobe_transform == 0) &&
                     !(!J.adobe && J.comp[0].id == 'R' && J.comp[1].id == 'G' && J.comp[2].id == 'B');
    if (J.ncomp == 1 || (!want_color && ycc)) {
        Component &c = J.comp[0];
        J.idct_component(c);
        std::vector<uint8_t> full;
        int fw = 0;
        upsample_component(J, c, full, fw);
        channels = 1;
        pix.resize((size_t)w * h);
        for (int y = 0; y < h; ++y) memcpy(&pix[(size_t)y * w], &full[(size_t)y * fw], w);
        apply_orientation(J.orientation, w, h, 1, pix);
        return true;
    }
    std::vector<uint8_t> pl[3];
    int fw[3];
    for (int i = 0; i < 3; ++i) {
        J.idct_component(J.comp[i]);
        upsample_component(J, J.comp[i], pl[i], fw[i]);
    }
    channels = 3;
    pix.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int Y = pl[0][(size_t)y * fw[0] + x], cb = pl[1][(size_t)y * fw[1] + x], cr = pl[2][(size_t)y * fw[2] + x];
            uint8_t *o = &pix[((size_t)y * w + x) * 3];
            if (!ycc) { o[0] = (uint8_t)Y; o[1] = (uint8_t)cb; o[2] = (uint8_t)cr; continue; }
            // jdcolor.c build_ycc_rgb_table: 16-bit fixed point, ONE_HALF folded into the tables
            const int cbx = cb - 128, crx = cr - 128;
            const int r = Y + (int)((91881L * crx + 32768) >> 16);
            const int g = Y + (int)((-22554L * cbx - 46802L * crx + 32768) >> 16);
            const int b = Y + (int)((116130L * cbx + 32768) >> 16);
            o[0] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
            o[1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g);
            o[2] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b);
        }
    if (!want_color) {  // RGB-coded JPEG read as grey: cv2 converts BGR2GRAY after decoding
        std::vector<uint8_t> g((size_t)w * h);
        for (size_t i = 0; i < g.size(); ++i) g[i] = (uint8_t)((pix[3 * i] * 4899 + pix[3 * i + 1] * 9617 + pix[3 * i + 2] * 1868 + 8192) >> 14);
        pix.swap(g);
        channels = 1;
    }
    apply_orientation(J.orientation, w, h, channels, pix);
    return true;
}

}  // namespace apd

// ------------------------------------------------------------------------------------------------ encoder
// Baseline JPEG writer for the debug imagery of the drop-in (depth_k.jpg, normal_k.jpg: APD.cpp:162-230, 265-287): 8-bit,
// 4:4:4, the Annex K quantisation tables at OpenCV's default quality 95, and one flat prefix code per class (every DC
// category 4 bits, every AC (run, size) symbol 8 bits) declared in the DHT segments -- valid for any decoder, a little larger
// than with tuned tables, and free of transcribed 162-entry tables.
namespace apd {
namespace {

struct BitWriter {
    std::vector<uint8_t> &out;
    uint32_t acc = 0;
    int n = 0;
    explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
    void put(uint32_t code, int len) {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            out.push_back(b);
            if (b == 0xFF) out.push_back(0);
            n -= 8;
        }
    }
    void flush() { if (n) put(0x7F, 8 - n); }
};

const uint8_t kStdLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                               14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                                 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// AC symbols of baseline JPEG: EOB, ZRL and (run 0..15, size 1..10); flat 8-bit codes in this order
int ac_symbol_index(int sym) {
    if (sym == 0x00) return 0;
    if (sym == 0xF0) return 1;
    return 2 + (sym >> 4) * 10 + ((sym & 15) - 1);
}

}  // namespace

bool encode_jpeg(int w, int h, int channels, const uint8_t *pix /* grey or BGR */, int quality, std::vector<uint8_t> &out) {
    if (w <= 0 || h <= 0 || (channels != 1 && channels != 3) || w > 65535 || h > 65535) return false;
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    uint8_t q[2][64];
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) {
            const int v = ((t ? kStdChromaQ[i] : kStdLumaQ[i]) * scale + 50) / 100;
            q[t][i] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
        }
    auto be16 = [&](int v) { out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); };
    out.clear();
    out.push_back(0xFF); out.push_back(0xD8);
    static const uint8_t jfif[] = {0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    out.insert(out.end(), jfif, jfif + sizeof(jfif));
    for (int t = 0; t < (channels == 3 ? 2 : 1); ++t) {
        out.push_back(0xFF); out.push_back(0xDB); be16(67); out.push_back((uint8_t)t);
        for (int i = 0; i < 64; ++i) out.push_back(q[t][kZigzag[i]]);
    }
    out.push_back(0xFF); out.push_back(0xC0); be16(8 + 3 * channels); out.push_back(8); be16(h); be16(w); out.push_back((uint8_t)channels);
    for (int c = 0; c < channels; ++c) { out.push_back((uint8_t)(c + 1)); out.push_back(0x11); out.push_back((uint8_t)(c ? 1 : 0)); }
    // DHT: DC table 0 = 12 categories with 4-bit codes 0..11; AC table 0 = 162 symbols with 8-bit codes 0..161
    out.push_back(0xFF); out.push_back(0xC4); be16(2 + 1 + 16 + 12); out.push_back(0x00);
    for (int l = 1; l <= 16; ++l) out.push_back((uint8_t)(l == 4 ? 12 : 0));
    for (int i = 0; i < 12; ++i) out.push_back((uint8_t)i);
    out.push_back(0xFF); out.push_back(0xC4); be16(2 + 1 + 16 + 162); out.push_back(0x10);
    for (int l = 1; l <= 16; ++l) out.push_back((uint8_t)(l == 8 ? 162 : 0));
    out.push_back(0x00); out.push_back(0xF0);
    for (int r = 0; r < 16; ++r) for (int s = 1; s <= 10; ++s) out.push_back((uint8_t)((r << 4) | s));
    out.push_back(0xFF); out.push_back(0xDA); be16(6 + 2 * channels); out.push_back((uint8_t)channels);
    for (int c = 0; c < channels; ++c) { out.push_back((uint8_t)(c + 1)); out.push_back(0x00); }
    out.push_back(0); out.push_back(63); out.push_back(0);

    float ct[8][8];
    for (int u = 0; u < 8; ++u)
        for (int x = 0; x < 8; ++x) ct[u][x] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    BitWriter bw(out);
    int pred[3] = {0, 0, 0};
    for (int by = 0; by < h; by += 8)
        for (int bx = 0; bx < w; bx += 8)
            for (int c = 0; c < channels; ++c) {
                float blk[64], tmp[64];
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) {
                        const int sy = by + y < h ? by + y : h - 1, sx = bx + x < w ? bx + x : w - 1;
                        const uint8_t *p = pix + ((size_t)sy * w + sx) * channels;
                        float v;
                        if (channels == 1) v = p[0];
                        else {
                            const float B = p[0], G = p[1], R = p[2];
                            v = c == 0 ? 0.299f * R + 0.587f * G + 0.114f * B
                                       : c == 1 ? -0.168736f * R - 0.331264f * G + 0.5f * B + 128.0f : 0.5f * R - 0.418688f * G - 0.081312f * B + 128.0f;
                        }
                        blk[y * 8 + x] = v - 128.0f;
                    }
                for (int y = 0; y < 8; ++y)
                    for (int u = 0; u < 8; ++u) { float s = 0; for (int x = 0; x < 8; ++x) s += ct[u][x] * blk[y * 8 + x]; tmp[y * 8 + u] = s; }
                for (int u = 0; u < 8; ++u)
                    for (int v = 0; v < 8; ++v) { float s = 0; for (int y = 0; y < 8; ++y) s += ct[v][y] * tmp[y * 8 + u]; blk[v * 8 + u] = s; }
                int zz[64];
                const uint8_t *qt = q[c ? 1 : 0];
                for (int i = 0; i < 64; ++i) zz[i] = (int)std::lrint(blk[kZigzag[i]] / qt[kZigzag[i]]);
                auto category = [](int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; };
                auto bits_of = [](int v, int n) { return (uint32_t)(v < 0 ? v + (1 << n) - 1 : v); };
                const int diff = zz[0] - pred[c];
                pred[c] = zz[0];
                int n = category(diff);
                bw.put((uint32_t)n, 4);
                if (n) bw.put(bits_of(diff, n), n);
                int run = 0;
                for (int i = 1; i < 64; ++i) {
                    int v = zz[i];
                    if (v > 1023) v = 1023;
                    if (v < -1023) v = -1023;
                    if (v == 0) { ++run; continue; }
                    while (run > 15) { bw.put((uint32_t)ac_symbol_index(0xF0), 8); run -= 16; }
                    n = category(v);
                    bw.put((uint32_t)ac_symbol_index((run << 4) | n), 8);
                    bw.put(bits_of(v, n), n);
                    run = 0;
                }
                if (run) bw.put((uint32_t)ac_symbol_index(0x00), 8);
            }
    bw.flush();
    out.push_back(0xFF); out.push_back(0xD9);
    return true;
}

}  // namespace apd
