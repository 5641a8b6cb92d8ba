/*
 * tools/wide_bvh_model.cpp — ANALYSIS ONLY (CPU), groundwork for a 4-wide traversal layout.
 *
 * Collapses the reference's binary LinearBVHNode[] (bvh.cpp:223-245) into 4-wide nodes — every interior node N takes the
 * children of its interior children (a leaf child stays as it is): 2..4 children per wide node — and traverses them in
 * EXACTLY the order the reference's binary traversal visits the same subtrees (trace_bvh.cl:99-211: near child by the ray
 * sign on the split axis first, far child deferred, inclusive slab test, later equal-t hit wins, back-face culling), with
 * the same IEEE arithmetic (no FMA: built with -ffp-contract=off).  All children of a wide node are slab-tested when the
 * node is visited; deferred children carry their entry distance and are re-tested against the current t_max when popped.
 * Skipping the box test of the collapsed intermediate node relies on interval containment (child box inside parent
 * box => child interval inside parent interval under monotone rounding), which holds unless 0 * inf produces a NaN, i.e.
 * for rays whose direction has no zero component; other rays are reported as "literal" and not traversed here.
 *
 * wide_trace() returns per ray the primitive id, t and barycentrics; tools/wide_bvh_model.py compares them bit for bit
 * with the oracle's binary traversal and reports the step counts of both.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "rt_math.h"
#include "rt_types.h"

namespace
{
struct V3 { float x, y, z; };
inline V3 v3(const RtFloat3& f) { return V3{ f.x, f.y, f.z }; }
inline V3 operator-(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, V3 b) { return V3{ a.x * b.x, a.y * b.y, a.z * b.z }; }
inline V3 cross(V3 a, V3 b) { return V3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct WideNode
{
    // children in "reference order for an all-positive ray": slot 0,1 = children of the first binary child (or the child
    // itself in slot 0 when it is a leaf), slot 2,3 = same for the second binary child
    V3 bmin[4], bmax[4];
    int ref[4];            // >= 0: wide node index; < 0: ~binary leaf node index; INT32_MIN: empty slot
    uint8_t axis_n, axis_l, axis_r;   // split axes of N and of its two children (leaf children: unused)
    uint8_t l_interior, r_interior;
};
const int EMPTY = INT32_MIN;

struct Model
{
    const RtLinearBVHNode* nodes; const RtTriangle* tris;
    std::vector<WideNode> wide;
    std::vector<int> wide_of;       // binary interior node -> wide node
    int root_ref;
};

int collapse(Model& m, uint32_t n)
{
    const RtLinearBVHNode& N = m.nodes[n];
    int idx = (int)m.wide.size();
    m.wide.push_back(WideNode());
    m.wide_of[n] = idx;
    uint32_t kids[2] = { n + 1, N.offset };
    WideNode w; memset(&w, 0, sizeof(w));
    for (int s = 0; s < 4; ++s) w.ref[s] = EMPTY;
    w.axis_n = (uint8_t)(N.num_primitives_axis & 0xFFFF);
    for (int c = 0; c < 2; ++c)
    {
        const RtLinearBVHNode& K = m.nodes[kids[c]];
        bool leaf = (K.num_primitives_axis >> 16) > 0;
        if (c == 0) w.l_interior = !leaf; else w.r_interior = !leaf;
        if (leaf)
        {
            w.bmin[2 * c] = v3(K.bounds_min); w.bmax[2 * c] = v3(K.bounds_max); w.ref[2 * c] = ~(int)kids[c];
        }
        else
        {
            if (c == 0) w.axis_l = (uint8_t)(K.num_primitives_axis & 0xFFFF); else w.axis_r = (uint8_t)(K.num_primitives_axis & 0xFFFF);
            uint32_t g[2] = { kids[c] + 1, K.offset };
            for (int j = 0; j < 2; ++j)
            {
                const RtLinearBVHNode& G = m.nodes[g[j]];
                w.bmin[2 * c + j] = v3(G.bounds_min); w.bmax[2 * c + j] = v3(G.bounds_max);
                bool gleaf = (G.num_primitives_axis >> 16) > 0;
                w.ref[2 * c + j] = gleaf ? ~(int)g[j] : collapse(m, g[j]);
            }
        }
    }
    m.wide[idx] = w;
    return idx;
}

inline bool slab(V3 bmin, V3 bmax, V3 o, V3 inv, float t_min, float t_max, float* lo_out)
{
    V3 t0 = (bmin - o) * inv, t1 = (bmax - o) * inv;
    float lo = rt_fmaxf(rt_fmaxf(rt_fminf(t0.x, t1.x), rt_fminf(t0.y, t1.y)), rt_fminf(t0.z, t1.z));
    float hi = rt_fminf(rt_fminf(rt_fmaxf(t0.x, t1.x), rt_fmaxf(t0.y, t1.y)), rt_fmaxf(t0.z, t1.z));
    float tmin = rt_fmaxf(lo, t_min), tmax = rt_fminf(hi, t_max);
    *lo_out = tmin;
    return tmax >= tmin;
}

inline bool ray_triangle(V3 o, V3 d, float t_min, float t_max, V3 p1, V3 p2, V3 p3, float* u_out, float* v_out, float* t_out)
{
    V3 e1 = p2 - p1, e2 = p3 - p1;
    V3 pvec = cross(d, e2);
    float det = dot(e1, pvec);
    if (det < 1e-8f || -det > 1e-8f) return false;
    float inv_det = 1.0f / det;
    V3 tvec = o - p1;
    float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    V3 qvec = cross(tvec, e1);
    float v = dot(d, qvec) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = dot(e2, qvec) * inv_det;
    if (t < t_min || t > t_max) return false;
    *u_out = u; *v_out = v; *t_out = t;
    return true;
}
} // namespace

extern "C" {

/* hits[i] = {bc.x, bc.y, primitive id, t}; status[i]: 0 traced, 1 "literal" (zero / non-finite direction component: not traced);
 * counts[0] += wide nodes visited, [1] += child boxes tested, [2] += triangles tested, [3] += stack pops that were discarded */
void wide_trace(const RtLinearBVHNode* nodes, uint32_t n_nodes, const RtTriangle* tris, const RtRay* rays, uint32_t n_rays, int any_hit,
                RtHit* hits, uint8_t* status, uint64_t* counts)
{
    Model m; m.nodes = nodes; m.tris = tris; m.wide_of.assign(n_nodes, -1);
    bool root_leaf = (nodes[0].num_primitives_axis >> 16) > 0;
    m.root_ref = root_leaf ? ~0 : collapse(m, 0);
    uint64_t c_nodes = 0, c_boxes = 0, c_tris = 0, c_discard = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c_nodes, c_boxes, c_tris, c_discard)
    for (long long i = 0; i < (long long)n_rays; ++i)
    {
        V3 o = v3(rays[i].origin), d = v3(rays[i].direction);
        float t_min = rays[i].origin.w, t_max = rays[i].direction.w;
        RtHit h; h.bc.x = h.bc.y = 0.0f; h.t = 0.0f; h.primitive_id = RT_INVALID_ID;
        bool ok = std::isfinite(o.x) && std::isfinite(o.y) && std::isfinite(o.z) && std::isfinite(d.x) && std::isfinite(d.y) && std::isfinite(d.z) &&
                  d.x != 0.0f && d.y != 0.0f && d.z != 0.0f;
        status[i] = ok ? 0 : 1;
        if (!ok) { hits[i] = h; continue; }
        V3 inv = V3{ 1.0f / d.x, 1.0f / d.y, 1.0f / d.z };
        int sign[3] = { inv.x < 0, inv.y < 0, inv.z < 0 };
        float lo;
        if (!slab(v3(nodes[0].bounds_min), v3(nodes[0].bounds_max), o, inv, t_min, t_max, &lo)) { hits[i] = h; continue; }   // the root is tested like any node
        int stack_ref[64]; float stack_lo[64]; int sp = 0;
        int cur = m.root_ref;
        bool done = false;
        while (!done)
        {
            if (cur >= 0)
            {
                const WideNode& w = m.wide[cur];
                ++c_nodes;
                // reference visiting order of the four slots for this ray
                int g0 = sign[w.axis_n] ? 1 : 0;                       // group visited first: 0 = first binary child, 1 = second
                int order[4]; int k = 0;
                for (int gi = 0; gi < 2; ++gi)
                {
                    int g = gi == 0 ? g0 : 1 - g0;
                    bool interior = g == 0 ? w.l_interior : w.r_interior;
                    if (!interior) { order[k++] = 2 * g; continue; }
                    int first = sign[g == 0 ? w.axis_l : w.axis_r] ? 1 : 0;
                    order[k++] = 2 * g + first; order[k++] = 2 * g + (1 - first);
                }
                bool hit[4]; float clo[4];
                for (int j = 0; j < k; ++j)
                {
                    int s = order[j];
                    ++c_boxes;
                    hit[j] = slab(w.bmin[s], w.bmax[s], o, inv, t_min, t_max, &clo[j]);
                }
                int next = EMPTY; float next_lo = 0.0f; (void)next_lo;
                for (int j = k - 1; j >= 0; --j)                      // push in reverse, keep the first hit child as `next`
                {
                    if (!hit[j]) continue;
                    if (next != EMPTY) { stack_ref[sp] = next; stack_lo[sp] = next_lo; ++sp; }
                    next = w.ref[order[j]]; next_lo = clo[j];
                }
                if (next != EMPTY) { cur = next; continue; }
            }
            else
            {   // binary leaf node ~cur
                const RtLinearBVHNode& L = nodes[~cur];
                uint32_t np = L.num_primitives_axis >> 16;
                for (uint32_t p = 0; p < np; ++p)
                {
                    const RtTriangle& t = tris[L.offset + p];
                    ++c_tris;
                    float u, v, tt;
                    if (ray_triangle(o, d, t_min, t_max, v3(t.v1.position), v3(t.v2.position), v3(t.v3.position), &u, &v, &tt))
                    {
                        h.bc.x = u; h.bc.y = v; h.t = tt; h.primitive_id = L.offset + p; t_max = tt;
                        if (any_hit) { h.primitive_id = 0; done = true; break; }
                    }
                }
                if (done) break;
            }
            // pop: deferred children are re-tested against the (possibly shrunk) t_max
            done = true;
            while (sp > 0)
            {
                --sp;
                if (t_max >= stack_lo[sp]) { cur = stack_ref[sp]; done = false; break; }
                ++c_discard;
            }
        }
        hits[i] = h;
    }
    counts[0] += c_nodes; counts[1] += c_boxes; counts[2] += c_tris; counts[3] += c_discard;
}

} // extern "C"
