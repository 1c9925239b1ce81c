/*
 * jpeg_oracle.c -- CPU ORACLE (test infrastructure): a plain-C restatement of the baseline JPEG decode that the reference's
 * frames go through before the line path:
 *     duckietown_utils/jpg.py:21-31  image_cv_from_jpg -> cv2.imdecode(np.fromstring(data), cv2.IMREAD_COLOR)
 *     line_detector_node.py:155      image_cv = image_cv_from_jpg(image_msg.data)
 * cv2 decodes with libjpeg-turbo at its defaults: Huffman decode -> dequantise -> jpeg_idct_islow (13-bit constants) ->
 * "fancy" (triangle) chroma upsampling -> YCbCr->RGB with the 16-bit fixed-point tables.  The arithmetic of that pipeline is
 * integer and published (IJG / libjpeg-turbo jidctint.c, jdsample.c, jdcolor.c); it is restated here and pinned by
 * tests/test_oracle.py: bit-identical to cv2.imdecode on every real JPEG of the reference (tests/golden/real_images.npz)
 * and on re-encoded synthetic frames (4:2:0, 4:2:2, 4:4:4, grayscale, restart intervals, odd sizes).
 * Scope: baseline sequential DCT (SOF0 / SOF1 Huffman, 8-bit), 1 or 3 components, sampling factors 1 or 2.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    uint8_t bits[17], vals[256];
    int maxcode[18], valptr[17], mincode[17];
    int present;
} Huff;

typedef struct {
    int id, h, v, tq, td, ta;
    int bw, bh;            /* blocks per row / column of the component (padded to whole MCUs) */
    int dw, dh;            /* downsampled width / height in samples (real, unpadded) */
    int16_t *coef;         /* [bh][bw][64] natural order, dequantised on the fly */
    uint8_t *plane;        /* [bh*8][bw*8] */
    int pred;
} Comp;

typedef struct {
    const uint8_t *p, *end;
    uint32_t acc; int nbits; int hit_marker;
} Bits;

static const uint8_t ZZ[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static void huff_build(Huff *h)
{
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h->valptr[l] = k;
        h->mincode[l] = code;
        k += h->bits[l];
        code += h->bits[l];
        h->maxcode[l] = h->bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    h->maxcode[17] = 0x7fffffff;
    h->present = 1;
}

static void fill(Bits *b)
{
    while (b->nbits <= 24) {
        int c = 0;
        if (!b->hit_marker && b->p < b->end) {
            c = *b->p++;
            if (c == 0xFF) {
                int c2 = b->p < b->end ? *b->p : 0xD9;
                if (c2 == 0) ++b->p;                  /* stuffed zero */
                else { b->hit_marker = 1; --b->p; c = 0; }   /* a marker: feed zeros from here on */
            }
        }
        b->acc |= (uint32_t)c << (24 - b->nbits);
        b->nbits += 8;
    }
}
static int getbits(Bits *b, int n)
{
    if (n == 0) return 0;
    fill(b);
    int v = (int)(b->acc >> (32 - n));
    b->acc <<= n; b->nbits -= n;
    return v;
}
static int decode_sym(Bits *b, const Huff *h)
{
    fill(b);
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | (int)(b->acc >> 31);
        b->acc <<= 1; b->nbits -= 1;
        if (h->maxcode[l] >= 0 && code <= h->maxcode[l] && code >= h->mincode[l]) return h->vals[h->valptr[l] + code - h->mincode[l]];
        if (b->nbits == 0) fill(b);
    }
    return 0;
}
static int extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

/* jpeg_idct_islow (jidctint.c): CONST_BITS 13, PASS1_BITS 2 */
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static uint8_t range_limit_idct(int x)
{
    x &= 1023;                      /* RANGE_MASK */
    if (x >= 512) x -= 1024;
    x += 128;
    return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x);
}
static void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, int stride)
{
    int ws[64];
    for (int c = 0; c < 8; ++c) {
        int d0 = in[c] * q[c], d1 = in[8 + c] * q[8 + c], d2 = in[16 + c] * q[16 + c], d3 = in[24 + c] * q[24 + c];
        int d4 = in[32 + c] * q[32 + c], d5 = in[40 + c] * q[40 + c], d6 = in[48 + c] * q[48 + c], d7 = in[56 + c] * q[56 + c];
        int z2 = d2, z3 = d6;
        int z1 = (z2 + z3) * 4433;
        int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
        int tmp0 = (d0 + d4) * 8192, tmp1 = (d0 - d4) * 8192;
        int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; int z4 = tmp1 + tmp3;
        int z5 = (z3 + z4) * 9633;
        tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
        z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        ws[c] = DESCALE(tmp10 + tmp3, 11); ws[56 + c] = DESCALE(tmp10 - tmp3, 11);
        ws[8 + c] = DESCALE(tmp11 + tmp2, 11); ws[48 + c] = DESCALE(tmp11 - tmp2, 11);
        ws[16 + c] = DESCALE(tmp12 + tmp1, 11); ws[40 + c] = DESCALE(tmp12 - tmp1, 11);
        ws[24 + c] = DESCALE(tmp13 + tmp0, 11); ws[32 + c] = DESCALE(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; ++r) {
        const int *w = ws + 8 * r;
        int z2 = w[2], z3 = w[6];
        int z1 = (z2 + z3) * 4433;
        int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
        int tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
        int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; int z4 = tmp1 + tmp3;
        int z5 = (z3 + z4) * 9633;
        tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
        z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        uint8_t *o = out + r * stride;
        o[0] = range_limit_idct(DESCALE(tmp10 + tmp3, 18)); o[7] = range_limit_idct(DESCALE(tmp10 - tmp3, 18));
        o[1] = range_limit_idct(DESCALE(tmp11 + tmp2, 18)); o[6] = range_limit_idct(DESCALE(tmp11 - tmp2, 18));
        o[2] = range_limit_idct(DESCALE(tmp12 + tmp1, 18)); o[5] = range_limit_idct(DESCALE(tmp12 - tmp1, 18));
        o[3] = range_limit_idct(DESCALE(tmp13 + tmp0, 18)); o[4] = range_limit_idct(DESCALE(tmp13 - tmp0, 18));
    }
}

static int rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

/* fancy upsampling of one chroma plane to full resolution (jdsample.c: h2v2_fancy_upsample, h2v1_fancy_upsample; context
 * rows at the top / bottom edge are the edge row itself, jdmainct.c) */
static void upsample(const Comp *c, int hs, int vs, int W, int H, uint8_t *full /* [H][W] */)
{
    const int pw = c->bw * 8, dw = c->dw, dh = c->dh;
    if (hs == 1 && vs == 1) {
        for (int y = 0; y < H; ++y) memcpy(full + (size_t)y * W, c->plane + (size_t)y * pw, W);
        return;
    }
    uint8_t *row = (uint8_t *)malloc((size_t)dw * 2 + 2);
    for (int y = 0; y < H; ++y) {
        if (hs == 2 && vs == 2) {
            const int sy = y >> 1;
            int ny = (y & 1) ? sy + 1 : sy - 1;                 /* the nearer neighbour row */
            if (ny < 0) ny = 0;
            if (ny > dh - 1) ny = dh - 1;
            const uint8_t *in0 = c->plane + (size_t)sy * pw, *in1 = c->plane + (size_t)ny * pw;
            if (dw == 1) {
                int s = in0[0] * 3 + in1[0];
                row[0] = (uint8_t)((s * 4 + 8) >> 4); row[1] = (uint8_t)((s * 4 + 7) >> 4);
            } else {
                int this_ = in0[0] * 3 + in1[0], next = in0[1] * 3 + in1[1], last;
                row[0] = (uint8_t)((this_ * 4 + 8) >> 4);
                row[1] = (uint8_t)((this_ * 3 + next + 7) >> 4);
                last = this_; this_ = next;
                for (int x = 2; x < dw; ++x) {
                    next = in0[x] * 3 + in1[x];
                    row[2 * x - 2] = (uint8_t)((this_ * 3 + last + 8) >> 4);
                    row[2 * x - 1] = (uint8_t)((this_ * 3 + next + 7) >> 4);
                    last = this_; this_ = next;
                }
                row[2 * dw - 2] = (uint8_t)((this_ * 3 + last + 8) >> 4);
                row[2 * dw - 1] = (uint8_t)((this_ * 4 + 7) >> 4);
            }
        } else if (hs == 2 && vs == 1) {
            const uint8_t *in = c->plane + (size_t)y * pw;
            if (dw == 1) { row[0] = row[1] = in[0]; }
            else {
                row[0] = in[0];
                row[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
                for (int x = 1; x < dw - 1; ++x) {
                    row[2 * x] = (uint8_t)((in[x] * 3 + in[x - 1] + 1) >> 2);
                    row[2 * x + 1] = (uint8_t)((in[x] * 3 + in[x + 1] + 2) >> 2);
                }
                row[2 * dw - 2] = (uint8_t)((in[dw - 1] * 3 + in[dw - 2] + 1) >> 2);
                row[2 * dw - 1] = in[dw - 1];
            }
        } else {   /* h1v2: libjpeg-turbo >= 1.5 has h1v2_fancy_upsample; rows only */
            const int sy = y >> 1;
            int ny = (y & 1) ? sy + 1 : sy - 1;
            if (ny < 0) ny = 0;
            if (ny > dh - 1) ny = dh - 1;
            const uint8_t *in0 = c->plane + (size_t)sy * pw, *in1 = c->plane + (size_t)ny * pw;
            const int bias = (y & 1) ? 2 : 1;
            for (int x = 0; x < dw; ++x) row[x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
        }
        memcpy(full + (size_t)y * W, row, W);
    }
    free(row);
}

static uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* Header probe: returns 0 and the frame size / component count, or < 0 when the stream is not a baseline JPEG this
 * restatement covers. */
ORC_API int orc_jpeg_info(const uint8_t *data, size_t len, int *W, int *H, int *ncomp)
{
    size_t i = 2;
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return -1;
    while (i + 4 <= len) {
        if (data[i] != 0xFF) return -1;
        int m = data[i + 1];
        if (m == 0xFF) { ++i; continue; }
        int L = rd16(data + i + 2);
        if (m == 0xC0 || m == 0xC1) {
            if (data[i + 4] != 8) return -2;
            *H = rd16(data + i + 5); *W = rd16(data + i + 7); *ncomp = data[i + 9];
            return 0;
        }
        if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) return -2;   /* progressive / lossless / arithmetic */
        i += 2 + (size_t)L;
    }
    return -1;
}

/* Decode to BGR (3 components) or gray replicated to BGR (1 component), out = [H][W][3].  Returns 0 or < 0. */
ORC_API int orc_jpeg_decode_bgr(const uint8_t *data, size_t len, uint8_t *out)
{
    uint16_t qt[4][64];
    Huff dc[4], ac[4];
    Comp comp[3];
    int W = 0, H = 0, nc = 0, hmax = 1, vmax = 1, restart = 0, have_sof = 0;
    memset(dc, 0, sizeof(dc)); memset(ac, 0, sizeof(ac)); memset(comp, 0, sizeof(comp)); memset(qt, 0, sizeof(qt));
    size_t i = 2;
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return -1;
    int rc = -1;
    while (i + 4 <= len) {
        if (data[i] != 0xFF) goto done;
        int m = data[i + 1];
        if (m == 0xFF) { ++i; continue; }
        if (m == 0xD9) break;
        int L = rd16(data + i + 2);
        const uint8_t *s = data + i + 4, *e = data + i + 2 + L;
        if (e > data + len) goto done;
        if (m == 0xDB) {
            while (s < e) {
                int pq = s[0] >> 4, tq = s[0] & 15; ++s;
                if (tq > 3) goto done;
                for (int k = 0; k < 64; ++k) { qt[tq][ZZ[k]] = (uint16_t)(pq ? rd16(s) : s[0]); s += pq ? 2 : 1; }
            }
        } else if (m == 0xC4) {
            while (s < e) {
                int tc = s[0] >> 4, th = s[0] & 15; ++s;
                if (th > 3) goto done;
                Huff *h = tc ? &ac[th] : &dc[th];
                int n = 0;
                h->bits[0] = 0;
                for (int l = 1; l <= 16; ++l) { h->bits[l] = s[l - 1]; n += s[l - 1]; }
                s += 16;
                if (n > 256) goto done;
                memcpy(h->vals, s, n); s += n;
                huff_build(h);
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (s[0] != 8) { rc = -2; goto done; }
            H = rd16(s + 1); W = rd16(s + 3); nc = s[5];
            if ((nc != 1 && nc != 3) || W <= 0 || H <= 0) { rc = -2; goto done; }
            for (int c = 0; c < nc; ++c) {
                comp[c].id = s[6 + 3 * c]; comp[c].h = s[7 + 3 * c] >> 4; comp[c].v = s[7 + 3 * c] & 15; comp[c].tq = s[8 + 3 * c];
                if (comp[c].h < 1 || comp[c].h > 2 || comp[c].v < 1 || comp[c].v > 2 || comp[c].tq > 3) { rc = -2; goto done; }
                if (comp[c].h > hmax) hmax = comp[c].h;
                if (comp[c].v > vmax) vmax = comp[c].v;
            }
            if (nc == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }
            have_sof = 1;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            rc = -2; goto done;
        } else if (m == 0xDD) {
            restart = rd16(s);
        } else if (m == 0xDA) {
            if (!have_sof) goto done;
            int ns = s[0];
            if (ns != nc) { rc = -2; goto done; }
            for (int k = 0; k < ns; ++k) {
                int cid = s[1 + 2 * k], c = -1;
                for (int q = 0; q < nc; ++q) if (comp[q].id == cid) c = q;
                if (c < 0) goto done;
                comp[c].td = s[2 + 2 * k] >> 4; comp[c].ta = s[2 + 2 * k] & 15;
            }
            /* geometry */
            const int mcuw = 8 * hmax, mcuh = 8 * vmax, mx = (W + mcuw - 1) / mcuw, my = (H + mcuh - 1) / mcuh;
            for (int c = 0; c < nc; ++c) {
                Comp *cp = &comp[c];
                cp->bw = mx * cp->h; cp->bh = my * cp->v;
                cp->dw = (W * cp->h + hmax - 1) / hmax; cp->dh = (H * cp->v + vmax - 1) / vmax;
                cp->coef = (int16_t *)calloc((size_t)cp->bw * cp->bh * 64, sizeof(int16_t));
                cp->plane = (uint8_t *)malloc((size_t)cp->bw * cp->bh * 64);
                cp->pred = 0;
            }
            Bits b; b.p = e; b.end = data + len; b.acc = 0; b.nbits = 0; b.hit_marker = 0;
            int todo = restart;
            for (int y = 0; y < my; ++y)
                for (int x = 0; x < mx; ++x) {
                    if (restart && todo == 0) {
                        /* byte-align, expect RSTn */
                        b.acc = 0; b.nbits = 0; b.hit_marker = 0;
                        while (b.p + 1 < b.end && !(b.p[0] == 0xFF && b.p[1] >= 0xD0 && b.p[1] <= 0xD7)) ++b.p;
                        if (b.p + 1 < b.end) b.p += 2;
                        for (int c = 0; c < nc; ++c) comp[c].pred = 0;
                        todo = restart;
                    }
                    for (int c = 0; c < nc; ++c) {
                        Comp *cp = &comp[c];
                        for (int by = 0; by < cp->v; ++by)
                            for (int bx = 0; bx < cp->h; ++bx) {
                                int16_t *blk = cp->coef + ((size_t)(y * cp->v + by) * cp->bw + (x * cp->h + bx)) * 64;
                                int t = decode_sym(&b, &dc[cp->td]);
                                int diff = t ? extend(getbits(&b, t), t) : 0;
                                cp->pred += diff;
                                blk[0] = (int16_t)cp->pred;
                                for (int k = 1; k < 64;) {
                                    int rs = decode_sym(&b, &ac[cp->ta]);
                                    int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) {
                                        if (r != 15) break;
                                        k += 16;
                                        continue;
                                    }
                                    k += r;
                                    if (k > 63) break;
                                    blk[ZZ[k]] = (int16_t)extend(getbits(&b, sz), sz);
                                    ++k;
                                }
                            }
                    }
                    if (restart) --todo;
                }
            /* IDCT */
            for (int c = 0; c < nc; ++c) {
                Comp *cp = &comp[c];
                for (int by = 0; by < cp->bh; ++by)
                    for (int bx = 0; bx < cp->bw; ++bx)
                        idct_islow(cp->coef + ((size_t)by * cp->bw + bx) * 64, qt[cp->tq], cp->plane + ((size_t)by * 8) * (cp->bw * 8) + bx * 8, cp->bw * 8);
            }
            /* upsample + colour */
            if (nc == 1) {
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        uint8_t v = comp[0].plane[(size_t)y * comp[0].bw * 8 + x];
                        uint8_t *o = out + ((size_t)y * W + x) * 3;
                        o[0] = o[1] = o[2] = v;
                    }
            } else {
                uint8_t *full[3];
                for (int c = 0; c < 3; ++c) {
                    full[c] = (uint8_t *)malloc((size_t)W * H);
                    upsample(&comp[c], hmax / comp[c].h, vmax / comp[c].v, W, H, full[c]);
                }
                /* jdcolor.c build_ycc_rgb_table / ycc_rgb_convert */
                int crr[256], cbb[256], crg[256], cbg[256];
                for (int k = 0; k < 256; ++k) {
                    int x = k - 128;
                    crr[k] = (91881 * x + 32768) >> 16;      /* FIX(1.40200) */
                    cbb[k] = (116130 * x + 32768) >> 16;     /* FIX(1.77200) */
                    crg[k] = -46802 * x;                     /* FIX(0.71414) */
                    cbg[k] = -22554 * x + 32768;             /* FIX(0.34414) */
                }
                for (size_t p = 0; p < (size_t)W * H; ++p) {
                    int yy = full[0][p], cb = full[1][p], cr = full[2][p];
                    out[3 * p + 2] = clamp8(yy + crr[cr]);
                    out[3 * p + 1] = clamp8(yy + ((cbg[cb] + crg[cr]) >> 16));
                    out[3 * p + 0] = clamp8(yy + cbb[cb]);
                }
                for (int c = 0; c < 3; ++c) free(full[c]);
            }
            rc = 0;
            goto done;
        }
        i += 2 + (size_t)L;
    }
done:
    for (int c = 0; c < 3; ++c) { free(comp[c].coef); free(comp[c].plane); }
    return rc;
}
