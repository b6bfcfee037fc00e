// jpeg2ppm -- decodes a baseline JPEG with host/jpeg_decode.hpp and writes the pixels as binary PPM/PGM (test helper
// for the ingest step; also a quick way to pre-convert images).   usage: jpeg2ppm in.jpg out.ppm
#include <cstdio>

#include "jpeg_decode.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) { printf("usage: %s in.jpg out.ppm\n", argv[0]); return 2; }
    int w, h, c;
    std::vector<uint8_t> px;
    const std::string err = jpegdec::load_jpeg(argv[1], w, h, c, px);
    if (!err.empty()) { printf("Error: %s\n", err.c_str()); return 1; }
    FILE *fp = fopen(argv[2], "wb");
    if (!fp) { printf("Error: cannot write %s\n", argv[2]); return 1; }
    fprintf(fp, "%s\n%d %d\n255\n", c == 1 ? "P5" : "P6", w, h);
    fwrite(px.data(), 1, px.size(), fp);
    fclose(fp);
    printf("%dx%d, %d channel(s)\n", w, h, c);
    return 0;
}
