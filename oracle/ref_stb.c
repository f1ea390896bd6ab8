/* ref_stb.c — decodes an image file with the reference's OWN image decoder (third-party/stb_image.h, compiled where it lies) exactly
 * as src/HostScene.cpp:10-51 calls it for an embedded texture (stbi_load_from_memory, req_comp 0) — or, with a third argument 3, as it calls
 * it for a texture FILE (stbi_load(path, ..., 3), :29) — and dumps "W H C\n" + the raw bytes (C = the channels of the dump).
 * TEST INFRASTRUCTURE (oracle/): pins multi-gpu-path-tracer_b200/csrc/host/JpegDecoder.h, BmpTgaDecoder.h (and the PNG reader) to the reference's texels. */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_stb <image> <out.raw> [req_comp]\n"); return 2; }
    int req = argc > 3 ? atoi(argv[3]) : 0;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *buf = (unsigned char *)malloc((size_t)n);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) return 1;
    fclose(f);
    int w, h, c;
    unsigned char *px = stbi_load_from_memory(buf, (int)n, &w, &h, &c, req);
    if (req) c = req;
    if (!px) { fprintf(stderr, "stb_image: %s\n", stbi_failure_reason()); return 3; }
    FILE *o = fopen(argv[2], "wb");
    fprintf(o, "%d %d %d\n", w, h, c);
    fwrite(px, 1, (size_t)w * h * c, o);
    fclose(o);
    return 0;
}
