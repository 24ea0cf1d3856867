/* Minimal baseline-TIFF reader exporting the seven libtiff symbols GMA.c uses.
 *
 * TEST INFRASTRUCTURE (oracle build only).  Supports what the synthetic
 * generators in this repo write: little-endian classic TIFF, one sample per
 * pixel, 8 or 16 bits, uncompressed, any strip layout.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tiffio.h"

struct mini_tiff {
    FILE *f;
    uint32_t width, height, bits, spp, rows_per_strip, nstrips;
    uint32_t *strip_off;
};

static uint32_t rd_val(FILE *f, uint16_t type, const unsigned char *field) {
    (void)f;
    if (type == 3) return (uint32_t)(field[0] | (field[1] << 8));
    return (uint32_t)field[0] | ((uint32_t)field[1] << 8) | ((uint32_t)field[2] << 16) | ((uint32_t)field[3] << 24);
}

TIFF *TIFFOpen(const char *name, const char *mode) {
    (void)mode;
    FILE *f = fopen(name, "rb");
    if (!f) return NULL;
    unsigned char hdr[8];
    if (fread(hdr, 1, 8, f) != 8 || hdr[0] != 'I' || hdr[1] != 'I' || hdr[2] != 42) { fclose(f); return NULL; }
    uint32_t ifd = rd_val(f, 4, hdr + 4);
    TIFF *t = (TIFF *)calloc(1, sizeof(TIFF));
    t->f = f; t->spp = 1; t->bits = 8; t->rows_per_strip = 0xffffffffu;
    fseek(f, ifd, SEEK_SET);
    unsigned char nb[2];
    if (fread(nb, 1, 2, f) != 2) { fclose(f); free(t); return NULL; }
    int nent = nb[0] | (nb[1] << 8);
    uint32_t so_count = 0, so_off = 0; uint16_t so_type = 4;
    for (int i = 0; i < nent; i++) {
        unsigned char e[12];
        if (fread(e, 1, 12, f) != 12) break;
        uint16_t tag = e[0] | (e[1] << 8), type = e[2] | (e[3] << 8);
        uint32_t count = rd_val(f, 4, e + 4);
        uint32_t v = rd_val(f, type, e + 8);
        switch (tag) {
            case 256: t->width = v; break;
            case 257: t->height = v; break;
            case 258: t->bits = v; break;
            case 277: t->spp = v; break;
            case 278: t->rows_per_strip = v; break;
            case 273: so_count = count; so_type = type; so_off = rd_val(f, 4, e + 8);
                      if (count == 1) so_off = v; break;
            default: break;
        }
    }
    if (t->rows_per_strip > t->height) t->rows_per_strip = t->height;
    t->nstrips = (t->height + t->rows_per_strip - 1) / t->rows_per_strip;
    t->strip_off = (uint32_t *)calloc(t->nstrips ? t->nstrips : 1, sizeof(uint32_t));
    if (so_count == 1) {
        t->strip_off[0] = so_off;
    } else {
        fseek(f, so_off, SEEK_SET);
        for (uint32_t i = 0; i < so_count && i < t->nstrips; i++) {
            unsigned char b[4] = {0, 0, 0, 0};
            if (fread(b, 1, so_type == 3 ? 2 : 4, f) == 0) break;
            t->strip_off[i] = rd_val(f, so_type, b);
        }
    }
    return t;
}

void TIFFClose(TIFF *t) { if (t) { fclose(t->f); free(t->strip_off); free(t); } }

int TIFFGetField(TIFF *t, uint32_t tag, ...) {
    va_list ap; va_start(ap, tag);
    uint32_t *p = va_arg(ap, uint32_t *);
    va_end(ap);
    switch (tag) {
        case TIFFTAG_IMAGEWIDTH: *p = t->width; return 1;
        case TIFFTAG_IMAGELENGTH: *p = t->height; return 1;
        case TIFFTAG_SAMPLESPERPIXEL: *p = t->spp; return 1;
        default: return 0; /* TIFFTAG_DATATYPE: absent, like modern libtiff */
    }
}

tsize_t TIFFScanlineSize(TIFF *t) { return (tsize_t)t->width * t->spp * (t->bits / 8); }

int TIFFReadScanline(TIFF *t, tdata_t buf, uint32_t row, uint16_t sample) {
    (void)sample;
    uint32_t strip = row / t->rows_per_strip, r = row % t->rows_per_strip;
    tsize_t ls = TIFFScanlineSize(t);
    fseek(t->f, (long)t->strip_off[strip] + (long)r * ls, SEEK_SET);
    return fread(buf, 1, (size_t)ls, t->f) == (size_t)ls ? 1 : -1;
}

void *_TIFFmalloc(tsize_t s) { return calloc(1, (size_t)s); }
void _TIFFfree(void *p) { free(p); }
