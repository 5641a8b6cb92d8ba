/*
 * obj_reader.cpp — see obj_reader.hpp.  Every rule below is one the reference's OBJ loader (tinyobjloader v2.0,
 * 3rdparty/tinyobjloader/tiny_obj_loader.h, cited as tol:line) applies and the OBJ format itself leaves open.
 */
#include "obj_reader.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>

namespace rt_host
{
namespace obj
{
namespace
{

inline bool is_space(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// Lines end with "\n", "\r\n" or a lone "\r"; a last line without an end still counts (tol:731-763)
class LineReader
{
public:
    explicit LineReader(std::string text) : text_(std::move(text)) {}
    bool next(std::string& line)
    {
        if (pos_ >= text_.size()) return false;
        size_t e = pos_;
        while (e < text_.size() && text_[e] != '\n' && text_[e] != '\r') ++e;
        line.assign(text_, pos_, e - pos_);
        if (e < text_.size()) e += (text_[e] == '\r' && e + 1 < text_.size() && text_[e + 1] == '\n') ? 2 : 1;
        pos_ = e;
        return true;
    }
private:
    std::string text_;
    size_t pos_ = 0;
};

bool read_text(const std::string& path, std::string& out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

// a number token ends at the next blank; what does not parse leaves the default (tol:951-959)
float parse_real(const char*& tok, double default_value = 0.0)
{
    tok += strspn(tok, " \t");
    const char* end = tok + strcspn(tok, " \t\r");
    double v = default_value;
    ParseDouble(tok, end, &v);
    tok = end;
    return (float)v;
}

void skip_tokens(const char*& tok, int n)
{
    for (int i = 0; i < n; ++i)
    {
        tok += strspn(tok, " \t");
        tok += strcspn(tok, " \t\r");
    }
}

std::string first_word(const char*& tok)
{
    tok += strspn(tok, " \t");
    size_t n = strcspn(tok, " \t\r");
    std::string s(tok, n);
    tok += n;
    return s;
}

bool has_key(const char* tok, const char* key)
{
    size_t n = strlen(key);
    return strncmp(tok, key, n) == 0 && is_space(tok[n]);
}

// "map_Xx [options] file name": each known option takes a fixed number of arguments (on/off and single numbers 1, -mm 2,
// -o / -s / -t always 3 — also when fewer are numbers), everything from the first other word to the end of the line is the file
// name, blanks included (tol:1191-1273).  false = options only.
bool texture_name(const char* tok, std::string& name)
{
    static const struct { const char* opt; int args; } options[] = {
        { "-blendu", 1 }, { "-blendv", 1 }, { "-clamp", 1 }, { "-boost", 1 }, { "-bm", 1 }, { "-o", 3 }, { "-s", 3 }, { "-t", 3 },
        { "-type", 1 }, { "-texres", 1 }, { "-imfchan", 1 }, { "-mm", 2 }, { "-colorspace", 1 },
    };
    bool found = false;
    while (!is_eol(*tok))
    {
        tok += strspn(tok, " \t");
        bool option = false;
        for (auto const& o : options)
            if (has_key(tok, o.opt))
            {
                tok += strlen(o.opt);
                skip_tokens(tok, o.args);
                option = true;
                break;
            }
        if (option) continue;
        name = tok;
        tok += name.size();
        found = true;
    }
    return found;
}

// tol:1823-2214.  The record under construction is pushed when the next "newmtl" arrives (only if it has a name) and once more
// at the end of the file (always): a file without "newmtl" yields one unnamed default material.  A name maps to the FIRST
// material that carries it.  "Kd seen" is never cleared, so the 0.6 grey a lone map_Kd implies is only given while no Kd line
// has been read in the whole file.
void read_mtl(const std::string& text, std::vector<Material>& materials, std::map<std::string, int>& by_name)
{
    Material m;
    bool has_kd = false;
    LineReader lines(text);
    std::string line;
    auto flush = [&]() {
        by_name.insert(std::make_pair(m.name, (int)materials.size()));
        materials.push_back(m);
    };
    while (lines.next(line))
    {
        size_t last = line.find_last_not_of(" \t");
        line.erase(last == std::string::npos ? 0 : last + 1);
        if (line.empty()) continue;
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (*tok == '\0' || *tok == '#') continue;
        auto read3 = [&](float* out) { tok += 2; out[0] = parse_real(tok); out[1] = parse_real(tok); out[2] = parse_real(tok); };
        if (has_key(tok, "newmtl"))
        {
            if (!m.name.empty()) flush();
            m = Material();
            m.name = tok + 7;
        }
        else if (has_key(tok, "Kd")) { read3(m.diffuse); has_kd = true; }
        else if (has_key(tok, "Ks")) read3(m.specular);
        else if (has_key(tok, "Kt") || has_key(tok, "Tf")) read3(m.transmittance);
        else if (has_key(tok, "Ni")) { tok += 2; m.ior = parse_real(tok); }
        else if (has_key(tok, "Ke")) read3(m.emission);
        else if (has_key(tok, "Pr")) { tok += 2; m.roughness = parse_real(tok); }
        else if (has_key(tok, "Pm")) { tok += 2; m.metallic = parse_real(tok); }
        else if (has_key(tok, "map_Kd"))
        {
            texture_name(tok + 7, m.diffuse_tex);
            if (!has_kd) m.diffuse[0] = m.diffuse[1] = m.diffuse[2] = 0.6f;
        }
        else if (has_key(tok, "map_Ks")) texture_name(tok + 7, m.specular_tex);
        else if (has_key(tok, "map_d")) { m.alpha_tex = tok + 6; texture_name(tok + 6, m.alpha_tex); }
        else if (has_key(tok, "map_Pr")) texture_name(tok + 7, m.roughness_tex);
        else if (has_key(tok, "map_Pm")) texture_name(tok + 7, m.metallic_tex);
        else if (has_key(tok, "map_Ke")) texture_name(tok + 7, m.emissive_tex);
    }
    flush();
}

// 1-based -> 0-based, negative = relative to the elements read so far; 0 is not an index (tol:771-791)
bool fix_index(int idx, int n, int* out)
{
    if (idx > 0) { *out = idx - 1; return true; }
    if (idx == 0) return false;
    *out = n + idx;
    return true;
}

// "v", "v/vt", "v//vn", "v/vt/vn" (tol:1105-1156)
bool parse_corner(const char*& tok, int nv, int nvn, int nvt, Index& out)
{
    Index c;
    if (!fix_index(atoi(tok), nv, &c.v)) return false;
    tok += strcspn(tok, "/ \t\r");
    if (*tok != '/') { out = c; return true; }
    ++tok;
    if (*tok == '/')
    {
        ++tok;
        if (!fix_index(atoi(tok), nvn, &c.vn)) return false;
        tok += strcspn(tok, "/ \t\r");
        out = c;
        return true;
    }
    if (!fix_index(atoi(tok), nvt, &c.vt)) return false;
    tok += strcspn(tok, "/ \t\r");
    if (*tok != '/') { out = c; return true; }
    ++tok;
    if (!fix_index(atoi(tok), nvn, &c.vn)) return false;
    tok += strcspn(tok, "/ \t\r");
    out = c;
    return true;
}

// crossing-number test of (tx, ty) against a triangle (tol:1355-1367)
bool inside_triangle(const float* vx, const float* vy, float tx, float ty)
{
    bool c = false;
    for (int i = 0, j = 2; i < 3; j = i++)
        if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
    return c;
}

// Faces are cut into triangles when their group is closed (material change, "g", "o", end of file), with the vertex positions
// read up to that point (tol:1370-1720):
//   3 corners: kept as is;
//   4 corners: split along the SHORTER diagonal — (0 1 2)(0 2 3) if |v2 - v0|^2 < |v3 - v1|^2, else (0 1 3)(1 2 3); dropped if a
//              corner names a vertex that does not exist;
//   more:      ear clipping in the plane of the two axes chosen from the first non-degenerate corner, in single precision.
struct Triangulator
{
    const std::vector<float>& v;
    Mesh& out;

    bool valid(int vi) const { return (size_t)3 * (size_t)vi + 2 < v.size(); }
    void emit(const Index& a, const Index& b, const Index& c, int material)
    {
        out.indices.push_back(a); out.indices.push_back(b); out.indices.push_back(c);
        out.material_ids.push_back(material);
    }
    void face(const std::vector<Index>& f, int material)
    {
        size_t n = f.size();
        if (n < 3) return;
        if (n == 3) { emit(f[0], f[1], f[2], material); return; }
        if (n == 4)
        {
            for (int k = 0; k < 4; ++k)
                if (!valid(f[k].v)) return;
            const float* p0 = &v[(size_t)f[0].v * 3]; const float* p1 = &v[(size_t)f[1].v * 3];
            const float* p2 = &v[(size_t)f[2].v * 3]; const float* p3 = &v[(size_t)f[3].v * 3];
            float e02x = p2[0] - p0[0], e02y = p2[1] - p0[1], e02z = p2[2] - p0[2];
            float e13x = p3[0] - p1[0], e13y = p3[1] - p1[1], e13z = p3[2] - p1[2];
            float sqr02 = e02x * e02x + e02y * e02y + e02z * e02z;
            float sqr13 = e13x * e13x + e13y * e13y + e13z * e13z;
            if (sqr02 < sqr13) { emit(f[0], f[1], f[2], material); emit(f[0], f[2], f[3], material); }
            else { emit(f[0], f[1], f[3], material); emit(f[1], f[2], f[3], material); }
            return;
        }
        polygon(f, material);
    }
    void polygon(const std::vector<Index>& f, int material)
    {
        size_t n = f.size();
        // projection axes: drop the axis along which the first real corner's normal is largest
        size_t axes[2] = { 1, 2 };
        for (size_t k = 0; k < n; ++k)
        {
            int i0 = f[k % n].v, i1 = f[(k + 1) % n].v, i2 = f[(k + 2) % n].v;
            if (!valid(i0) || !valid(i1) || !valid(i2)) continue;
            const float* a = &v[(size_t)i0 * 3]; const float* b = &v[(size_t)i1 * 3]; const float* c = &v[(size_t)i2 * 3];
            float e0x = b[0] - a[0], e0y = b[1] - a[1], e0z = b[2] - a[2];
            float e1x = c[0] - b[0], e1y = c[1] - b[1], e1z = c[2] - b[2];
            float cx = std::fabs(e0y * e1z - e0z * e1y);
            float cy = std::fabs(e0z * e1x - e0x * e1z);
            float cz = std::fabs(e0x * e1y - e0y * e1x);
            const float eps = std::numeric_limits<float>::epsilon();
            if (cx > eps || cy > eps || cz > eps)
            {
                if (!(cx > cy && cx > cz))
                {
                    axes[0] = 0;
                    if (cz > cx && cz > cy) axes[1] = 1;
                }
                break;
            }
        }
        auto coord_ok = [&](int vi, size_t axis) { return (size_t)vi * 3 + axis < v.size(); };
        float area = 0;
        for (size_t k = 0; k < n; ++k)
        {
            int i0 = f[k % n].v, i1 = f[(k + 1) % n].v;
            if (!coord_ok(i0, axes[0]) || !coord_ok(i0, axes[1]) || !coord_ok(i1, axes[0]) || !coord_ok(i1, axes[1])) continue;
            float x0 = v[(size_t)i0 * 3 + axes[0]], y0 = v[(size_t)i0 * 3 + axes[1]];
            float x1 = v[(size_t)i1 * 3 + axes[0]], y1 = v[(size_t)i1 * 3 + axes[1]];
            area += (x0 * y1 - y0 * x1) * 0.5f;
        }
        std::vector<Index> rest = f;
        size_t guess = 0;
        size_t budget = n, previous = n;          // rounds left without removing a corner
        while (rest.size() > 3 && budget > 0)
        {
            size_t m = rest.size();
            if (guess >= m) guess -= m;
            if (previous != m) { previous = m; budget = m; } else --budget;
            Index ind[3];
            float vx[3], vy[3];
            for (size_t k = 0; k < 3; ++k)
            {
                ind[k] = rest[(guess + k) % m];
                int vi = ind[k].v;
                if (!coord_ok(vi, axes[0]) || !coord_ok(vi, axes[1])) { vx[k] = 0.0f; vy[k] = 0.0f; }
                else { vx[k] = v[(size_t)vi * 3 + axes[0]]; vy[k] = v[(size_t)vi * 3 + axes[1]]; }
            }
            float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0], e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
            float cross = e0x * e1y - e0y * e1x;
            if (cross * area < 0.0f) { ++guess; continue; }         // a reflex corner
            bool overlap = false;
            for (size_t other = 3; other < m; ++other)
            {
                size_t idx = (guess + other) % m;
                int ovi = rest[idx].v;
                if (!coord_ok(ovi, axes[0]) || !coord_ok(ovi, axes[1])) continue;
                if (inside_triangle(vx, vy, v[(size_t)ovi * 3 + axes[0]], v[(size_t)ovi * 3 + axes[1]])) { overlap = true; break; }
            }
            if (overlap) { ++guess; continue; }
            emit(ind[0], ind[1], ind[2], material);                 // an ear: cut it off at its middle corner
            rest.erase(rest.begin() + (std::ptrdiff_t)((guess + 1) % m));
        }
        if (rest.size() == 3) emit(rest[0], rest[1], rest[2], material);
    }
};

} // namespace

// Decimal string -> double the way the reference's reader does it (tol:837-949): digits are accumulated into a double
// (integer part: m = m * 10 + d; fraction: m += d * 10^-k with the first seven powers as literals and pow(10, -k) beyond), a
// decimal exponent e is applied as ldexp(m * pow(5, e), e).  Not correctly rounded — the double differs from strtod's for more
// than half of all 6+ digit fractions, and for strings next to a float rounding boundary ("1.9662156701087952") so does the float
// the renderer gets; strtod is therefore not a substitute.
bool ParseDouble(const char* s, const char* end, double* result)
{
    if (s >= end) return false;
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', exp_sign = '+';
    const char* p = s;
    bool leading_dot = false;
    if (*p == '+' || *p == '-')
    {
        sign = *p++;
        if (p != end && *p == '.') leading_dot = true;
    }
    else if (is_digit(*p)) {}
    else if (*p == '.') leading_dot = true;
    else return false;

    int read = 0;
    if (!leading_dot)
    {
        while (p != end && is_digit(*p)) { mantissa *= 10; mantissa += (int)(*p - '0'); ++p; ++read; }
        if (read == 0) return false;
    }
    bool have_exponent = false;
    if (p != end)
    {
        if (*p == '.')
        {
            static const double negative_powers[] = { 1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001 };
            ++p;
            read = 1;
            while (p != end && is_digit(*p))
            {
                mantissa += (int)(*p - '0') * (read < 8 ? negative_powers[read] : std::pow(10.0, -read));
                ++read; ++p;
            }
            have_exponent = p != end && (*p == 'e' || *p == 'E');
        }
        else have_exponent = *p == 'e' || *p == 'E';
    }
    if (have_exponent)
    {
        ++p;
        if (p != end && (*p == '+' || *p == '-')) exp_sign = *p++;
        else if (p != end && is_digit(*p)) {}
        else return false;
        read = 0;
        while (p != end && is_digit(*p))
        {
            if (exponent > std::numeric_limits<int>::max() / 10) return false;
            exponent = exponent * 10 + (int)(*p - '0');
            ++p; ++read;
        }
        if (exp_sign == '-') exponent = -exponent;
        if (read == 0) return false;
    }
    *result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    return true;
}

bool Read(const char* filename, const std::string& mtl_dir, Mesh& out, std::string& error)
{
    std::string text;
    if (!read_text(filename, text)) { error = std::string("Cannot open file [") + filename + "]"; return false; }
    out = Mesh();
    std::map<std::string, int> material_by_name;
    std::vector<std::vector<Index>> open_faces;       // faces of the group being read, not yet cut into triangles
    int material = -1;
    Triangulator tri{ out.positions, out };
    auto close_group = [&]() {
        for (auto const& f : open_faces) tri.face(f, material);
        open_faces.clear();
    };

    LineReader lines(std::move(text));
    std::string line;
    size_t line_no = 0;
    while (lines.next(line))
    {
        ++line_no;
        if (line.empty()) continue;
        const char* tok = line.c_str();
        tok += strspn(tok, " \t");
        if (*tok == '\0' || *tok == '#') continue;

        if (tok[0] == 'v' && is_space(tok[1]))
        {   // x y z (a colour may follow; not used)
            tok += 2;
            for (int i = 0; i < 3; ++i) out.positions.push_back(parse_real(tok));
        }
        else if (tok[0] == 'v' && tok[1] == 'n' && is_space(tok[2]))
        {
            tok += 3;
            for (int i = 0; i < 3; ++i) out.normals.push_back(parse_real(tok));
        }
        else if (tok[0] == 'v' && tok[1] == 't' && is_space(tok[2]))
        {
            tok += 3;
            for (int i = 0; i < 2; ++i) out.texcoords.push_back(parse_real(tok));
        }
        else if (tok[0] == 'f' && is_space(tok[1]))
        {
            tok += 2;
            tok += strspn(tok, " \t");
            std::vector<Index> face;
            face.reserve(4);
            while (!is_eol(*tok))
            {
                Index c;
                if (!parse_corner(tok, (int)(out.positions.size() / 3), (int)(out.normals.size() / 3), (int)(out.texcoords.size() / 2), c))
                {
                    error = "Failed parse `f' line (e.g. zero value for face index), line " + std::to_string(line_no);
                    return false;
                }
                face.push_back(c);
                tok += strspn(tok, " \t\r");
            }
            open_faces.push_back(std::move(face));
        }
        else if (strncmp(tok, "usemtl", 6) == 0)
        {   // the name is the first word after the key; an unknown name selects "no material"
            tok += 6;
            std::string name = first_word(tok);
            auto it = material_by_name.find(name);
            int next = it == material_by_name.end() ? -1 : it->second;
            if (next != material) { close_group(); material = next; }
        }
        else if (has_key(tok, "mtllib"))
        {   // blank-separated file names ('\' escapes a blank): the first one that opens is read, the others are ignored
            tok += 7;
            std::vector<std::string> names;
            std::string cur;
            bool escaping = false;
            for (const char* q = tok; *q; ++q)
            {
                if (escaping) escaping = false;
                else if (*q == '\\') { escaping = true; continue; }
                else if (*q == ' ') { if (!cur.empty()) names.push_back(cur); cur.clear(); continue; }
                cur += *q;
            }
            names.push_back(cur);
            for (auto const& n : names)
            {
                std::string path = mtl_dir.empty() ? n : (mtl_dir.back() == '/' ? mtl_dir + n : mtl_dir + "/" + n);
                std::string mtl_text;
                if (read_text(path, mtl_text)) { read_mtl(mtl_text, out.materials, material_by_name); break; }
            }
        }
        else if ((tok[0] == 'g' || tok[0] == 'o') && is_space(tok[1])) close_group();
    }
    close_group();
    return true;
}

} // namespace obj
} // namespace rt_host
