// tools/image_fuzz.cpp — mutation fuzz of the texture decoders (host/image_loader.cpp, host/jpeg_decoder.cpp) under
// AddressSanitizer + UBSan: every input file is damaged a few hundred times (byte flips, random bytes, truncation) and decoded;
// the decoders must refuse or decode, never read or write out of bounds.
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Iinclude -Iraytracing_b200/host tools/image_fuzz.cpp \
//       raytracing_b200/host/jpeg_decoder.cpp raytracing_b200/host/image_loader.cpp -lz -o /tmp/image_fuzz
//   /tmp/image_fuzz tests/golden/textures/*.jpg tests/golden/textures/*.png
#include "reference_api.hpp"

#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <unistd.h>

int main(int argc, char** argv)
{
    using namespace rt_host;
    std::mt19937 rng(5);
    long decoded = 0, refused = 0;
    char tmpl[] = "/tmp/image_fuzz_XXXXXX";
    if (!mkdtemp(tmpl)) return 1;
    for (int a = 1; a < argc; ++a)
    {
        FILE* f = fopen(argv[a], "rb");
        if (!f) continue;
        std::vector<unsigned char> d(1 << 22);
        d.resize(fread(d.data(), 1, d.size(), f));
        fclose(f);
        const char* ext = strrchr(argv[a], '.');
        std::string path = std::string(tmpl) + "/x" + (ext ? ext : ".png");
        for (int it = 0; it < 300 && !d.empty(); ++it)
        {
            std::vector<unsigned char> m = d;
            int nmut = 1 + rng() % 4;
            for (int k = 0; k < nmut; ++k)
            {
                size_t pos = rng() % m.size();
                switch (rng() % 3)
                {
                case 0: m[pos] = (unsigned char)rng(); break;
                case 1: m[pos] ^= (unsigned char)(1u << (rng() % 8)); break;
                case 2: m.resize(pos + 1); break;
                }
            }
            FILE* o = fopen(path.c_str(), "wb");
            fwrite(m.data(), 1, m.size(), o);
            fclose(o);
            TextureImage img; std::string err;
            if (LoadTextureImage(path.c_str(), img, err)) ++decoded; else ++refused;
        }
        unlink(path.c_str());
    }
    rmdir(tmpl);
    printf("decoded %ld, refused %ld\n", decoded, refused);
    return 0;
}
