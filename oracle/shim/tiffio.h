/* Declaration-only stand-in for libtiff's <tiffio.h>.
 *
 * TEST INFRASTRUCTURE (oracle build only).  The reference's GMA.h:23-25 includes
 * "tiffio.h" unconditionally, but libtiff-dev is not installed in this image.
 * This header declares exactly the seven libtiff entry points and the types/tags
 * GMA.c:246-316 uses, so the reference translation units compile unmodified.
 * The symbols are provided by minitiff.c (uncompressed strip TIFF only).
 */
#ifndef _TIFFIO_
#define _TIFFIO_
#include <stdint.h>
#include <stddef.h>

typedef struct mini_tiff TIFF;
typedef int64_t tsize_t;
typedef void *tdata_t;
typedef uint32_t uint32;
typedef uint16_t uint16;
typedef uint8_t uint8;

#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_SAMPLESPERPIXEL 277
#define TIFFTAG_DATATYPE 32996

TIFF *TIFFOpen(const char *name, const char *mode);
void TIFFClose(TIFF *tif);
int TIFFGetField(TIFF *tif, uint32_t tag, ...);
tsize_t TIFFScanlineSize(TIFF *tif);
int TIFFReadScanline(TIFF *tif, tdata_t buf, uint32_t row, uint16_t sample);
void *_TIFFmalloc(tsize_t s);
void _TIFFfree(void *p);
#endif
