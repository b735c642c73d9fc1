// drr_scene.cpp -- host front-end of libdrr: the reference's Renderer (BSP walk, seg clipping, occlusion arrays, visplane
// building, sprite projection / clipping / ordering) restated in C++ so that it EMITS per-frame draw lists through the
// C ABI (include/drr.h) instead of drawing.  The reference is Rust and no Rust toolchain exists in this image, so this
// file plays the role of the "thin Rust host" of the north star; INTEGRATION.md shows the equivalent Rust-side change.
//
// Mirrors (reference paths relative to its repo root):
//   Game::new asset/map loading        src/game.rs:118-196, src/wad.rs, src/map/*.rs, src/graphics/*.rs, src/map_objects.rs:25-50
//   Renderer::render                   src/renderer/mod.rs:118-136
//   Segs::process_seg/process_sidedef  src/renderer/segs.rs:121-590
//   clip_to_viewport, projection       src/renderer/misc.rs:13-161
//   SidedefVisPlanes                   src/renderer/sidedef_visplanes.rs
//   draw_map_objects                   src/renderer/map_objects.rs:19-241
//   BitmapRender ordering predicates   src/renderer/bitmap_render.rs:137-188
//
// Unlike the reference everything name-keyed is resolved once at load time (textures, flats, sprite frames -> integer
// handles), per-frame state lives in reusable flat arrays, and nothing here touches a pixel.
//
// Arithmetic contract: compiled with -ffp-contract=off, no fast-math; f32 expressions keep the reference's evaluation
// order; `as` casts saturate; i16 arithmetic wraps; sinf/cosf/sqrtf/fmodf come from the host libm exactly like Rust's
// f32::{sin,cos,sqrt,%}.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/drr.h"
#include "../../../data/drr_info_table.inc"

namespace {

struct Panic {
    std::string msg;
};
[[noreturn]] void panic(const std::string &m) { throw Panic{m}; }

// ---- Rust scalar semantics ---------------------------------------------------------------------------------------
inline int16_t as_i16(float f) {
    if (f != f) return 0;
    if (f <= -32768.0f) return INT16_MIN;
    if (f >= 32767.0f) return INT16_MAX;
    return (int16_t)f;
}
inline int32_t as_i32(float f) {
    if (f != f) return 0;
    if (f <= -2147483648.0f) return INT32_MIN;
    if (f >= 2147483648.0f) return INT32_MAX;
    return (int32_t)f;
}
inline uint8_t as_u8(float f) {
    if (f != f || f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
inline uint64_t as_usize(float f) {
    if (f != f || f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}
inline int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
const float PI_F = 3.14159265358979323846f;

struct V2 {
    float x, y;
};
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
struct Seg2 {
    V2 s, e;
};
inline V2 rot(V2 v, float a) { return {v.x * cosf(a) - v.y * sinf(a), v.y * cosf(a) + v.x * sinf(a)}; } // vertexes.rs:20-25
inline float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
inline bool left_of(V2 v, const Seg2 &l) { return cross(v - l.s, l.e - l.s) <= 0.0f; } // vertexes.rs:32-34
inline float dist(V2 a, V2 b) {
    float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
bool intersect(const Seg2 &a, const Seg2 &b, V2 *out) { // geometry.rs:56-82
    float x1 = a.s.x, y1 = a.s.y, x2 = a.e.x, y2 = a.e.y, x3 = b.s.x, y3 = b.s.y, x4 = b.e.x, y4 = b.e.y;
    float quot = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (fabsf(quot) < 0.001f) return false;
    float inv = 1.0f / quot;
    out->x = inv * ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4));
    out->y = inv * ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4));
    return true;
}

// ---- WAD ---------------------------------------------------------------------------------------------------------
std::string up(std::string s) {
    for (auto &c : s)
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return s;
}
struct Lump {
    std::string name;
    uint32_t off, size;
};
struct Wad {
    std::vector<uint8_t> b;
    std::vector<Lump> lumps;
    std::map<std::string, int> by_name; // last duplicate wins (wad.rs:153-155)
    int s_start = -1, s_end = -1;
    uint8_t u8(size_t o) const {
        if (o >= b.size()) panic("WAD read past end of file");
        return b[o];
    }
    int16_t i16(size_t o) const { return (int16_t)(uint16_t)(u8(o) | (u8(o + 1) << 8)); }
    uint32_t u32(size_t o) const { return u8(o) | (u8(o + 1) << 8) | (u8(o + 2) << 16) | ((uint32_t)u8(o + 3) << 24); }
    float f16(size_t o) const { return (float)i16(o); }
    std::string name8(size_t o) const { // wad.rs:112-126
        std::string s;
        bool nul_terminated = u8(o + 7) == 0;
        for (int i = 0; i < 8; i++) {
            if (nul_terminated && u8(o + i) == 0) break;
            s.push_back((char)u8(o + i));
        }
        return s;
    }
    int find(const std::string &n) const {
        auto it = by_name.find(up(n));
        return it == by_name.end() ? -1 : it->second;
    }
    const Lump &map_lump(const std::string &map, int k) const { // wad.rs:175-183
        std::string m = up(map);
        for (size_t i = 0; i < lumps.size(); i++)
            if (lumps[i].name == m) {
                if (i + k >= lumps.size()) panic("map lumps missing");
                return lumps[i + k];
            }
        panic("Could not find map " + map);
    }
    void parse() { // wad.rs:86-157
        if (b.size() < 12 || memcmp(b.data(), "IWAD", 4) != 0) panic("Unhandled WAD file type");
        uint32_t n = u32(4), dir = u32(8);
        for (uint32_t i = 0; i < n; i++) {
            size_t e = (size_t)dir + 16 * (size_t)i;
            Lump l{up(name8(e + 8)), u32(e), u32(e + 4)};
            by_name[l.name] = (int)lumps.size();
            lumps.push_back(l);
        }
        s_start = find("S_START");
        s_end = find("S_END");
        if (s_start < 0 || s_end < 0) panic("S_START / S_END missing");
    }
};

// ---- assets --------------------------------------------------------------------------------------------------------
struct HostBitmap {
    int16_t w = 0, h = 0;
    std::vector<int16_t> px; // row-major, -1 = None
};
struct HostPicture {
    int bitmap = -1;
    int16_t top_offset = 0;
};
struct SpriteFrame {
    bool rotate = false;
    std::vector<HostPicture> pics;
};

// ---- map -----------------------------------------------------------------------------------------------------------
struct FlatRef { // a sector flat resolved at load time: either one flat or an animation cycle (flats.rs:103-111)
    std::vector<int> ids; // flat handles; -2 = lump missing (panics if ever shown)
    std::vector<uint8_t> is_sky;
};
struct SectorH {
    int16_t floor, ceil, light;
    int16_t special = 0, light0 = 0; // sector special (light effects, thinkers.rs:14-76) and the WAD's light level (tic 0)
    FlatRef floor_flat, ceil_flat;
    bool ceil_name_has_sky; // segs.rs:464-469 tests the SECTOR's texture name, not the animated frame's
};
struct SideH {
    float xoff, yoff;
    int upper, lower, middle; // bitmap handle, -1 = "-", -2 = unknown texture name (reference panics when it is needed)
    std::string upper_n, lower_n, middle_n;
    int sector;
};
struct LineH {
    int16_t flags;
    int front, back;
};
struct SegH {
    V2 v1, v2;
    int line;
    bool dir;
    int16_t offset;
};
struct NodeH {
    float x, y, dx, dy;
    int right, left; // >= 0 node index, < 0: ~subsector
};
struct ObjH {
    int sprite;
    uint8_t frame;
    bool full_bright, is_null;
    V2 pos;
    float angle;
    int16_t state = 0, spawn_state = 0; // StateId now / at tic 0 (info.rs STATES)
};

// ---- per-frame records ---------------------------------------------------------------------------------------------
enum RState : uint8_t { SOLID, TWOSIDED, DRAWN, MAPOBJ };
struct Render { // BitmapRender, bitmap_render.rs:29-45
    RState state;
    int bitmap; // -1 none
    int16_t light;
    Seg2 line;
    float start_offset;
    int32_t sx, ex;
    float bottom_h, top_h;
    int16_t off_x, off_y;
    bool ext_bottom, ext_top, draw_ceiling;
    uint32_t col0, ncol; // range in Scene::colpool
};
struct PlaneH { // Visplane, visplanes.rs:17-26 ; arrays live in Scene::plane_rows
    int flat;   // flat handle
    bool sky;
    int16_t height, light, left, right;
    uint32_t rows; // offset of top[W] then bottom[W] in plane_rows
};

} // namespace

struct drr_scene {
    int W = 0, H = 0;
    float ASPECT, GCFX, CFX, CFY;
    std::string err;
    Wad wad;
    uint8_t palette[768];
    std::vector<HostBitmap> bitmaps;
    std::vector<std::array<uint8_t, 4096>> flats;
    std::vector<std::string> flat_names;
    std::map<std::string, int> flat_by_name, tex_by_name;
    std::map<std::string, std::vector<std::string>> animated;
    struct TexDef {
        int16_t w, h;
        std::vector<std::array<int16_t, 3>> patches;
    };
    std::map<std::string, TexDef> texdefs;
    std::vector<std::string> pnames;
    std::map<int, std::map<int, SpriteFrame>> sprites;
    int sky_bitmap = -1;

    std::vector<V2> verts;
    std::vector<SectorH> sectors;
    std::vector<SideH> sides;
    std::vector<LineH> lines;
    std::vector<SegH> segs;
    std::vector<std::pair<int, int>> ssectors; // first seg, count
    std::vector<NodeH> nodes;
    std::vector<ObjH> objects;
    bool have_start = false;
    float start[3];

    // what drr_scene_emit_views_device last uploaded, and where (the tables only depend on the flat animation frame and on
    // whether things are wanted)
    uint32_t fe_map_id = 0;
    const void *fe_ctx = nullptr;
    uint64_t fe_anim = ~0ull, fe_world = ~0ull;
    bool fe_things = false;

    // per-frame scratch, reused
    std::vector<Render> renders;
    std::vector<drr_col> colpool;
    std::vector<PlaneH> planes;
    std::vector<int16_t> plane_rows;
    std::vector<uint8_t> hor_ocl;
    std::vector<int16_t> floor_ocl, ceil_ocl, top_clip, bottom_clip;
    std::vector<Render> mo_renders;
    std::vector<drr_col> mo_cols;

    // ================================================================================================================
    // loading
    // ================================================================================================================
    int add_bitmap(HostBitmap &&b) {
        bitmaps.push_back(std::move(b));
        return (int)bitmaps.size() - 1;
    }
    bool decode_picture(const std::string &name, HostBitmap *bm, int16_t *top_offset) { // pictures.rs:66-126
        int li = wad.find(name);
        if (li < 0) return false;
        size_t off = wad.lumps[li].off;
        bm->w = wad.i16(off);
        bm->h = wad.i16(off + 2);
        if (bm->w < 0 || bm->h < 0) panic("negative picture size");
        if (top_offset) *top_offset = wad.i16(off + 6);
        bm->px.assign((size_t)bm->w * bm->h, -1);
        for (int c = 0; c < bm->w; c++) {
            size_t co = off + wad.u32(off + 8 + 4 * (size_t)c);
            for (;;) {
                uint8_t yo = wad.u8(co);
                if (yo == 0xff) break;
                uint8_t len = wad.u8(co + 1);
                for (int r = 0; r < len; r++) {
                    int y = r + yo;
                    if (y >= bm->h) panic("picture post outside the picture");
                    bm->px[(size_t)y * bm->w + c] = wad.u8(co + 3 + r);
                }
                co += (size_t)len + 4;
            }
        }
        return true;
    }
    int texture(const std::string &name) { // textures.rs:154-179 + Texture::load :74-103 ; -2 if unknown
        std::string key = up(name);
        auto hit = tex_by_name.find(key);
        if (hit != tex_by_name.end()) return hit->second;
        auto it = texdefs.find(key);
        if (it == texdefs.end()) return tex_by_name[key] = -2;
        const TexDef &td = it->second;
        HostBitmap bm;
        bm.w = td.w;
        bm.h = td.h;
        if (bm.w < 0 || bm.h < 0) panic("negative texture size");
        bm.px.assign((size_t)bm.w * bm.h, -1);
        for (auto &p : td.patches) {
            if (p[2] < 0 || (size_t)p[2] >= pnames.size()) panic("patch number out of range");
            HostBitmap pb;
            if (!decode_picture(pnames[p[2]], &pb, nullptr)) panic("missing patch lump " + pnames[p[2]]);
            for (int x = 0; x < pb.w; x++)
                for (int y = 0; y < pb.h; y++) {
                    int16_t px = wrap16(x + p[0]), py = wrap16(y + p[1]);
                    if (px >= 0 && px < bm.w && py >= 0 && py < bm.h)
                        bm.px[(size_t)py * bm.w + px] = pb.px[(size_t)y * pb.w + x]; // None overwrites too (textures.rs:97-98)
                }
        }
        return tex_by_name[key] = add_bitmap(std::move(bm));
    }
    int flat(const std::string &name) { // flats.rs:92-100,116-137 ; -2 if the lump is missing
        auto hit = flat_by_name.find(name);
        if (hit != flat_by_name.end()) return hit->second;
        int li = wad.find(name);
        if (li < 0) return flat_by_name[name] = -2;
        std::array<uint8_t, 4096> px;
        for (int i = 0; i < 4096; i++) px[i] = wad.u8(wad.lumps[li].off + i);
        flats.push_back(px);
        flat_names.push_back(name);
        return flat_by_name[name] = (int)flats.size() - 1;
    }
    FlatRef flat_ref(const std::string &name) {
        FlatRef r;
        auto it = animated.find(name);
        std::vector<std::string> names = it != animated.end() ? it->second : std::vector<std::string>{name};
        for (auto &n : names) {
            r.ids.push_back(flat(n));
            r.is_sky.push_back(n.find("SKY") != std::string::npos); // visplanes.rs:89 tests the drawn flat's own name
        }
        return r;
    }

    void load(const char *path, const char *map_name, int w, int h) {
        W = w;
        H = h;
        if (W <= 0 || H <= 0 || W > 32767 || H > 32767) panic("bad screen size");
        ASPECT = 200.0f / 240.0f; // constants.rs:7-17
        const float gsw = (float)(uint32_t)W / ASPECT;
        GCFX = gsw / 2.0f;
        CFX = (float)(uint32_t)W / 2.0f;
        CFY = (float)(uint32_t)H / 2.0f;

        FILE *f = fopen(path, "rb");
        if (!f) panic(std::string("cannot open ") + path);
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) wad.b.insert(wad.b.end(), buf, buf + n);
        fclose(f);
        wad.parse();

        int pp = wad.find("PLAYPAL"); // palette.rs:11-28
        if (pp < 0) panic("PLAYPAL missing");
        for (int i = 0; i < 768; i++) palette[i] = wad.u8(wad.lumps[pp].off + i);

        static const char *const anim[][5] = {{"NUKAGE1", "NUKAGE2", "NUKAGE3", nullptr}, {"FWATER1", "FWATER2", "FWATER3", "FWATER4", nullptr},
                                              {"SWATER1", "SWATER2", "SWATER3", "SWATER4", nullptr}, {"LAVA1", "LAVA2", "LAVA3", "LAVA4", nullptr},
                                              {"BLOOD1", "BLOOD2", "BLOOD3", nullptr}, {"RROCK05", "RROCK06", "RROCK07", "RROCK08", nullptr},
                                              {"SLIME01", "SLIME02", "SLIME03", "SLIME04", nullptr}, {"SLIME05", "SLIME06", "SLIME07", "SLIME08", nullptr},
                                              {"SLIME09", "SLIME10", "SLIME11", "SLIME12", nullptr}}; // flats.rs:30-75
        for (auto &row : anim) {
            std::vector<std::string> l;
            for (int i = 0; i < 5 && row[i]; i++) l.push_back(row[i]);
            for (auto &nm : l) animated[nm] = l;
        }

        // PNAMES / TEXTURE1 / TEXTURE2 (textures.rs:131-255)
        int pn = wad.find("PNAMES");
        if (pn < 0) panic("PNAMES missing");
        uint32_t np = wad.u32(wad.lumps[pn].off);
        for (uint32_t i = 0; i < np; i++) pnames.push_back(wad.name8(wad.lumps[pn].off + 4 + 8 * (size_t)i));
        for (const char *tl : {"TEXTURE1", "TEXTURE2"}) {
            int ti = wad.find(tl);
            if (ti < 0) {
                if (tl[7] == '1') panic("TEXTURE1 missing");
                continue;
            }
            size_t base = wad.lumps[ti].off;
            uint32_t cnt = wad.u32(base);
            for (uint32_t i = 0; i < cnt; i++) {
                size_t o = base + wad.u32(base + 4 + 4 * (size_t)i);
                TexDef td;
                td.w = wad.i16(o + 12);
                td.h = wad.i16(o + 14);
                int16_t pc = wad.i16(o + 20);
                for (int j = 0; j < pc; j++) {
                    size_t po = o + 22 + 10 * (size_t)j;
                    td.patches.push_back({wad.i16(po), wad.i16(po + 2), wad.i16(po + 4)});
                }
                texdefs[up(wad.name8(o))] = td;
            }
        }
        sky_bitmap = texture(sky_name(map_name)); // game.rs:199-227
        if (sky_bitmap < 0) panic("sky texture missing");

        load_map(map_name);
        load_sprites();
    }

    static std::string sky_name(const std::string &m) { // game.rs:199-227: regexes e(\d+)m(\d+) then (\d\d), unanchored
        for (size_t i = 0; i + 1 < m.size(); i++) {
            if (m[i] != 'e') continue;
            size_t j = i + 1;
            while (j < m.size() && isdigit((unsigned char)m[j])) j++;
            if (j == i + 1 || j + 1 >= m.size() || m[j] != 'm' || !isdigit((unsigned char)m[j + 1])) continue;
            long ep = strtol(m.substr(i + 1, j - i - 1).c_str(), nullptr, 10);
            return ep == 2 ? "SKY2" : ep == 3 ? "SKY3" : "SKY1";
        }
        for (size_t i = 0; i + 1 < m.size(); i++)
            if (isdigit((unsigned char)m[i]) && isdigit((unsigned char)m[i + 1])) {
                int k = (m[i] - '0') * 10 + (m[i + 1] - '0');
                return k < 12 ? "SKY1" : k < 21 ? "SKY2" : "SKY3";
            }
        return "SKY1";
    }

    void load_map(const std::string &name) {
        auto need = [](bool ok, const char *what) {
            if (!ok) panic(std::string("map index out of range: ") + what);
        };
        const Lump &lv = wad.map_lump(name, 4);
        for (size_t i = 0; i < lv.size / 4; i++) verts.push_back({wad.f16(lv.off + 4 * i), wad.f16(lv.off + 4 * i + 2)});
        const Lump &lsec = wad.map_lump(name, 8);
        for (size_t i = 0; i < lsec.size / 26; i++) {
            size_t o = lsec.off + 26 * i;
            SectorH s;
            s.floor = wad.i16(o);
            s.ceil = wad.i16(o + 2);
            std::string fn = wad.name8(o + 4), cn = wad.name8(o + 12);
            s.floor_flat = flat_ref(fn);
            s.ceil_flat = flat_ref(cn);
            s.ceil_name_has_sky = cn.find("SKY") != std::string::npos;
            s.light = s.light0 = wad.i16(o + 20);
            s.special = wad.i16(o + 22);
            sectors.push_back(std::move(s));
        }
        const Lump &lsd = wad.map_lump(name, 3);
        for (size_t i = 0; i < lsd.size / 30; i++) {
            size_t o = lsd.off + 30 * i;
            SideH s;
            s.xoff = wad.f16(o);
            s.yoff = wad.f16(o + 2);
            s.upper_n = wad.name8(o + 4);
            s.lower_n = wad.name8(o + 12);
            s.middle_n = wad.name8(o + 20);
            s.upper = s.upper_n == "-" ? -1 : texture(s.upper_n);
            s.lower = s.lower_n == "-" ? -1 : texture(s.lower_n);
            s.middle = s.middle_n == "-" ? -1 : texture(s.middle_n);
            s.sector = wad.i16(o + 28);
            need(s.sector >= 0 && (size_t)s.sector < sectors.size(), "sidedef sector");
            sides.push_back(std::move(s));
        }
        const Lump &lld = wad.map_lump(name, 2);
        for (size_t i = 0; i < lld.size / 14; i++) {
            size_t o = lld.off + 14 * i;
            LineH l{wad.i16(o + 4), wad.i16(o + 10), wad.i16(o + 12)};
            need(l.front >= -1 && l.front < (int)sides.size() && l.back >= -1 && l.back < (int)sides.size(), "linedef sidedef");
            need(wad.i16(o) >= 0 && (size_t)wad.i16(o) < verts.size() && wad.i16(o + 2) >= 0 && (size_t)wad.i16(o + 2) < verts.size(), "linedef vertex");
            lines.push_back(l);
        }
        const Lump &lsg = wad.map_lump(name, 5);
        for (size_t i = 0; i < lsg.size / 12; i++) {
            size_t o = lsg.off + 12 * i;
            int a = wad.i16(o), b = wad.i16(o + 2), ld = wad.i16(o + 6);
            need(a >= 0 && (size_t)a < verts.size() && b >= 0 && (size_t)b < verts.size() && ld >= 0 && (size_t)ld < lines.size(), "seg");
            segs.push_back({verts[a], verts[b], ld, wad.i16(o + 8) != 0, wad.i16(o + 10)});
        }
        const Lump &lss = wad.map_lump(name, 6);
        for (size_t i = 0; i < lss.size / 4; i++) {
            int cnt = wad.i16(lss.off + 4 * i), first = wad.i16(lss.off + 4 * i + 2);
            need(cnt <= 0 || (first >= 0 && (size_t)(first + cnt) <= segs.size()), "subsector segs");
            ssectors.push_back({first, std::max(cnt, 0)});
        }
        const Lump &lnd = wad.map_lump(name, 7);
        for (size_t i = 0; i < lnd.size / 28; i++) {
            size_t o = lnd.off + 28 * i;
            auto child = [&](int16_t v) -> int { // nodes.rs:17-27
                int idx = v & 0x7fff;
                if (v & INT16_MIN) {
                    need((size_t)idx < ssectors.size(), "node subsector");
                    return ~idx;
                }
                need((size_t)idx < nodes.size(), "node child");
                return idx;
            };
            NodeH n{wad.f16(o), wad.f16(o + 2), wad.f16(o + 4), wad.f16(o + 6), 0, 0};
            n.right = child(wad.i16(o + 24));
            n.left = child(wad.i16(o + 26));
            nodes.push_back(n);
        }
        if (nodes.empty()) panic("map has no nodes");
        const Lump &lth = wad.map_lump(name, 1);
        for (size_t i = 0; i < lth.size / 10; i++) {
            size_t o = lth.off + 10 * i;
            float x = wad.f16(o), y = wad.f16(o + 2), ang = wad.f16(o + 4) * (PI_F / 180.0f); // things.rs:36 to_radians
            int16_t type = wad.i16(o + 6);
            if (type == 1 && !have_start) {
                have_start = true;
                start[0] = x;
                start[1] = y;
                start[2] = ang;
            }
            if ((type >= 1 && type <= 4) || type == 11) continue; // map_objects.rs:31-36
            const DrrThingInfo *info = nullptr;
            for (auto &ti : DRR_THING_INFOS)
                if (ti.doomednum == type) info = &ti;
            if (!info) panic("unknown thing type " + std::to_string(type));
            const int16_t spawn = DRR_THING_SPAWN_STATE[info - DRR_THING_INFOS];
            objects.push_back({info->sprite, info->frame, info->full_bright != 0, info->is_null != 0, {x, y}, ang, spawn, spawn});
        }
    }

    void load_sprites() { // sprites.rs:26-97
        std::map<std::string, HostPicture> cache;
        auto get = [&](const std::string &lump) -> HostPicture {
            auto it = cache.find(lump);
            if (it != cache.end()) return it->second;
            HostBitmap bm;
            HostPicture p;
            if (!decode_picture(lump, &bm, &p.top_offset)) panic("sprite lump unreadable");
            p.bitmap = add_bitmap(std::move(bm));
            return cache[lump] = p;
        };
        for (int sid = 0; sid < 138; sid++) {
            const std::string prefix = DRR_SPRITE_NAMES[sid];
            std::map<int, std::map<int, HostPicture>> found;
            for (int li = wad.s_start; li < wad.s_end; li++) {
                const std::string &n = wad.lumps[li].name;
                if (n.compare(0, prefix.size(), prefix) != 0) continue;
                if (n.size() < 6 || n.size() == 7) panic("malformed sprite lump name " + n);
                HostPicture p = get(n);
                found[(uint8_t)(n[4] - 65)][(uint8_t)(n[5] - 48)] = p;
                if (n.size() > 6) { // second frame/rotation uses the mirrored picture (pictures.rs:129-147)
                    HostBitmap m = bitmaps[p.bitmap];
                    for (int y = 0; y < m.h; y++) std::reverse(m.px.begin() + (size_t)y * m.w, m.px.begin() + (size_t)(y + 1) * m.w);
                    HostPicture q = p;
                    q.bitmap = add_bitmap(std::move(m));
                    found[(uint8_t)(n[6] - 65)][(uint8_t)(n[7] - 48)] = q;
                }
            }
            for (auto &fr : found) {
                SpriteFrame sf;
                sf.rotate = fr.second.size() != 1;
                if (sf.rotate) {
                    if (fr.second.size() != 8) panic("Got something other than 8 rotations for " + prefix);
                    for (int r = 1; r <= 8; r++) {
                        if (!fr.second.count(r)) panic("sprite rotation missing");
                        sf.pics.push_back(fr.second[r]);
                    }
                } else {
                    if (!fr.second.count(0)) panic("single sprite rotation is not 0");
                    sf.pics.push_back(fr.second[0]);
                }
                sprites[sid][fr.first] = sf;
            }
        }
    }

    // ================================================================================================================
    // geometry helpers
    // ================================================================================================================
    struct ScreenLine {
        int32_t sx, sy, ex, ey;
    };
    ScreenLine project(const Seg2 &l, float height) const { // misc.rs:130-161
        V2 ts = {GCFX * l.s.y / l.s.x, GCFX * height / l.s.x};
        V2 te = {GCFX * l.e.y / l.e.x, GCFX * height / l.e.x};
        ts.x *= ASPECT;
        te.x *= ASPECT;
        ScreenLine r{as_i32(CFX - ts.x), as_i32(CFY - ts.y), as_i32(CFX - te.x), as_i32(CFY - te.y)};
        r.sx = std::min(r.sx, W - 1);
        r.ex = std::min(r.ex, W - 1);
        return r;
    }
    static bool clip_fov(const Seg2 &line, Seg2 *out, float *start_offset) { // misc.rs:13-115
        const Seg2 L = {{0.0f, 0.0f}, {1.0f, 1.0f}}, R = {{0.0f, 0.0f}, {1.0f, -1.0f}};
        bool s_out_l = left_of(line.s, L), e_out_l = left_of(line.e, L);
        bool s_out_r = !left_of(line.s, R), e_out_r = !left_of(line.e, R);
        bool s_in = line.s.x > 0.0f && !s_out_l && !s_out_r;
        bool e_in = line.e.x > 0.0f && !e_out_l && !e_out_r;
        if (s_in && e_in) {
            *out = line;
            *start_offset = 0.0f;
            return true;
        }
        V2 li, ri;
        bool lhit = intersect(line, L, &li) && li.x >= 0.0f;
        bool rhit = intersect(line, R, &ri) && ri.x >= 0.0f;
        if (!s_in && !e_in && !lhit && !rhit) return false;
        if (!s_in && !e_in && lhit != rhit) return false;
        if ((rhit && s_out_r && e_out_r) || (lhit && s_out_l && e_out_l)) return false;
        V2 s = line.s, e = line.e;
        float so = 0.0f;
        if (lhit) {
            if (s_out_l) {
                so = dist(li, s);
                s = li;
            }
            if (e_out_l) e = li;
        }
        if (rhit) {
            if (s_out_r) s = ri;
            if (e_out_r) e = ri;
        }
        *out = {s, e};
        *start_offset = so;
        return true;
    }
    int sector_at(V2 p) const { // renderer/bsp.rs:9-44
        int n = (int)nodes.size() - 1;
        for (;;) {
            const NodeH &nd = nodes[n];
            V2 a = {nd.x, nd.y};
            int ch = left_of(p, {a, a + V2{nd.dx, nd.dy}}) ? nd.left : nd.right;
            if (ch >= 0) {
                n = ch;
                continue;
            }
            auto ss = ssectors[~ch];
            for (int i = 0; i < ss.second; i++) {
                const SegH &sg = segs[ss.first + i];
                int sd = sg.dir ? lines[sg.line].back : lines[sg.line].front;
                if (sd != -1) return sides[sd].sector;
            }
            return -1;
        }
    }

    // ================================================================================================================
    // the time axis: the world `tic` game ticks after the start of the game (game.rs:456-482)
    // ================================================================================================================
    // Sector light effects (lights.rs) and map-object animation (map_objects.rs:63-95) are the two things that change
    // between tics and reach the draw path (sector light levels; sprite / frame / full-bright of the things).  The
    // reference seeds them from rand::thread_rng(); here ONE PCG32 stream per (seed) is consumed in the reference's thinker
    // order -- construction: sector effects in sector order, then the map objects; every tic: the same list order --
    // with gen_range(lo..hi) = lo + next % (hi - lo).  (include/drr.h: drr_scene_set_tic states the same contract.)
    uint32_t world_tic = 0;
    uint64_t world_seed = 0;
    uint64_t world_version = 0; // bumped by set_tic: the device front-end's map tables depend on it

    void set_tic(uint32_t tic, uint64_t seed) {
        uint64_t st = 0, inc = (0xda3e39cb94b95bdbull << 1) | 1u;
        auto next = [&]() -> uint32_t {
            const uint64_t old = st;
            st = old * 6364136223846793005ull + inc;
            const uint32_t x = (uint32_t)(((old >> 18u) ^ old) >> 27u), r = (uint32_t)(old >> 59u);
            return (x >> r) | (x << ((32u - r) & 31u));
        };
        next();
        st += seed;
        next();
        auto range = [&](int lo, int hi) { return (int16_t)(lo + (int)(next() % (uint32_t)(hi - lo))); };

        for (SectorH &c : sectors) c.light = c.light0;
        for (ObjH &o : objects) {
            o.state = o.spawn_state;
            apply_state(o);
        }
        struct Effect {
            int kind; // 1 flash, 2 strobe, 8 glow, 17 fire
            int sector;
            int16_t lo, hi, dark, count;
            bool up;
        };
        auto darkest_neighbour = [&](int sec, int16_t start) { // lights.rs:14-43
            int16_t m = start;
            for (const LineH &l : lines) {
                if (l.front == -1 || l.back == -1) continue;
                const int fs = sides[l.front].sector, bs = sides[l.back].sector;
                if (fs == sec) m = std::min(m, sectors[bs].light);
                if (bs == sec) m = std::min(m, sectors[fs].light);
            }
            return m;
        };
        std::vector<Effect> fx;
        for (int i = 0; i < (int)sectors.size(); i++) { // thinkers.rs:14-76
            const int16_t sp = sectors[i].special, lvl = sectors[i].light;
            if (sp == 1) {
                fx.push_back({1, i, darkest_neighbour(i, lvl), lvl, 0, range(1, 65), false}); // LightFlash::new: count in 1..=max_time(64)
            } else if (sp == 2 || sp == 3 || sp == 4 || sp == 12 || sp == 13) { // StrobeFlash::new
                int16_t lo = darkest_neighbour(i, lvl);
                if (lo == lvl) lo = 0;
                const bool sync = sp == 12 || sp == 13;
                const int16_t dark = (sp == 3 || sp == 12) ? 35 : 15; // SLOW_DARK / FAST_DARK
                fx.push_back({2, i, lo, lvl, dark, sync ? (int16_t)1 : range(1, 9), false});
            } else if (sp == 8) {
                fx.push_back({8, i, darkest_neighbour(i, lvl), lvl, 0, 0, false}); // GlowingLight::new
            } else if (sp == 17) {
                fx.push_back({17, i, (int16_t)(darkest_neighbour(i, lvl) + 16), lvl, 0, 4, false}); // FireFlicker::new
            }
        }
        std::vector<int16_t> mo_count(objects.size());
        for (size_t i = 0; i < objects.size(); i++) mo_count[i] = DRR_STATES[objects[i].state].tics; // MapObjectThinker::new

        for (uint32_t t = 0; t < tic; t++) {
            for (Effect &e : fx) {
                int16_t &lvl = sectors[e.sector].light;
                if (e.kind == 8) { // lights.rs:192-212, GLOW_SPEED = 8
                    if (e.up) {
                        lvl = wrap16(lvl + 8);
                        if (lvl >= e.hi) {
                            lvl = wrap16(lvl - 8);
                            e.up = false;
                        }
                    } else {
                        lvl = wrap16(lvl - 8);
                        if (lvl <= e.lo) {
                            lvl = wrap16(lvl + 8);
                            e.up = true;
                        }
                    }
                    continue;
                }
                e.count = wrap16(e.count - 1);
                if (e.count > 0) continue;
                if (e.kind == 1) { // lights.rs:81-101: min_time 7, max_time 64
                    if (lvl == e.hi) {
                        lvl = e.lo;
                        e.count = range(1, 8);
                    } else {
                        lvl = e.hi;
                        e.count = range(1, 65);
                    }
                } else if (e.kind == 2) { // lights.rs:144-164: STROBE_BRIGHT = 5
                    if (lvl == e.hi) {
                        lvl = e.lo;
                        e.count = e.dark;
                    } else {
                        lvl = e.hi;
                        e.count = 5;
                    }
                } else { // lights.rs:242-259
                    const int16_t amount = (int16_t)(range(0, 4) * 16);
                    lvl = wrap16(lvl - amount) < e.lo ? e.lo : wrap16(e.hi - amount);
                    e.count = 4;
                }
            }
            for (size_t i = 0; i < objects.size(); i++) { // map_objects.rs:84-95
                if (mo_count[i] == -1) continue;
                mo_count[i] = wrap16(mo_count[i] - 1);
                if (mo_count[i] > 0) continue;
                objects[i].state = DRR_STATES[objects[i].state].next;
                apply_state(objects[i]);
                mo_count[i] = DRR_STATES[objects[i].state].tics;
            }
        }
        world_tic = tic;
        world_seed = seed;
        world_version++;
    }
    static void apply_state(ObjH &o) { // MapObject.state -> what the renderer reads (renderer/map_objects.rs:37-63)
        const DrrState &st = DRR_STATES[o.state];
        o.sprite = st.sprite;
        o.frame = st.frame;
        o.full_bright = st.full_bright != 0;
        o.is_null = o.state == 0; // S_NULL
    }

    // ================================================================================================================
    // per-frame front-end
    // ================================================================================================================
    drr_ctx *ctx = nullptr;
    drr_recorder *rec = nullptr; // when set, the frame is recorded off-context (worker threads of drr_scene_emit_views)
    V2 ppos;
    float pfloor, pangle, timestamp;
    int phases;

    void chk(int rc) {
        if (rc != DRR_OK) panic(std::string("libdrr: ") + drr_error_name(rc) + ": " + (rec ? drr_recorder_last_error(rec) : drr_last_error(ctx)));
    }
    int sink_frame_begin(int idx, const drr_view *v) { return rec ? drr_recorder_frame_begin(rec, idx, v) : drr_frame_begin(ctx, idx, v); }
    int sink_columns(const drr_seg_hdr *h, const drr_col *c, int n) { return rec ? drr_recorder_emit_columns(rec, h, c, n) : drr_emit_columns(ctx, h, c, n); }
    int sink_visplane(const drr_visplane_hdr *h, const int16_t *t, const int16_t *b) {
        return rec ? drr_recorder_emit_visplane(rec, h, t, b) : drr_emit_visplane(ctx, h, t, b);
    }
    int sink_frame_end() { return rec ? drr_recorder_frame_end(rec) : drr_frame_end(ctx); }
    int sink_frame_abort() { return rec ? drr_recorder_frame_abort(rec) : drr_frame_abort(ctx); }
    void emit(const Render &r, const drr_col *cols, int phase) {
        if (r.bitmap < 0 || r.ncol == 0) return;
        drr_seg_hdr h;
        h.bitmap_id = r.bitmap;
        h.light_level = r.light;
        h.phase = (int16_t)phase;
        h.line_start_x = r.line.s.x;
        h.line_start_y = r.line.s.y;
        h.line_end_x = r.line.e.x;
        h.line_end_y = r.line.e.y;
        h.start_offset = r.start_offset;
        h.start_x = r.sx;
        h.end_x = r.ex;
        h.bottom_height = r.bottom_h;
        h.top_height = r.top_h;
        h.offset_x = r.off_x;
        h.offset_y = r.off_y;
        chk(sink_columns(&h, cols, (int)r.ncol));
    }

    int pick_flat(const FlatRef &fr, bool *sky) const { // flats.rs:103-111
        size_t k = fr.ids.size() == 1 ? 0 : (size_t)(as_usize(timestamp * 3.0f) % fr.ids.size());
        if (fr.ids[k] < 0) panic("flat lump missing");
        *sky = fr.is_sky[k] != 0;
        return fr.ids[k];
    }

    struct PlaneAcc { // SidedefVisPlanes, sidedef_visplanes.rs
        drr_scene *sc;
        int flat[2];
        bool sky[2];
        int16_t height[2], light;
        int cur[2] = {-1, -1}; // index into sc->planes of the plane being accumulated (0 bottom, 1 top)
        void point(int which, int16_t x, int16_t top_y, int16_t bottom_y) { // add_bottom_point :60-71 / add_top_point :73-84
            if (cur[which] < 0) {
                PlaneH p{flat[which], sky[which], height[which], light, x, x, (uint32_t)sc->plane_rows.size()};
                sc->plane_rows.resize(sc->plane_rows.size() + 2 * (size_t)sc->W, 0); // top/bottom zero-initialised (visplanes.rs:36-37)
                sc->planes.push_back(p);
                cur[which] = (int)sc->planes.size() - 1;
            }
            PlaneH &p = sc->planes[cur[which]];
            p.right = x;
            sc->plane_rows[p.rows + x] = top_y;
            sc->plane_rows[p.rows + sc->W + x] = bottom_y;
        }
        // flush :41-58 pushes bottom then top.  Planes are appended here when their FIRST point arrives, so when both are
        // open their relative order may differ from the reference's push order; fix it at flush time.
        void flush() {
            if (cur[0] >= 0 && cur[1] >= 0 && cur[1] < cur[0]) std::swap(sc->planes[cur[0]], sc->planes[cur[1]]);
            cur[0] = cur[1] = -1;
        }
    };

    void occlude(int16_t x) { // segs.rs:113-117
        hor_ocl[x] = 1;
        floor_ocl[x] = (int16_t)H / 2;
        ceil_ocl[x] = (int16_t)H / 2;
    }

    // segs.rs:121-350
    void sidedef_part(const Seg2 &cl, float start_offset, const SideH &sd, int16_t seg_offset, const SectorH &sec, int floor_flat, bool floor_sky,
                      int ceil_flat, bool ceil_sky, float bottom_h, float top_h, int32_t offset_y, int tex, const std::string &tex_name,
                      bool only_occ, bool lower, bool upper, bool draw_ceiling, bool two_sided_mid) {
        ScreenLine bottom = project(cl, bottom_h), top = project(cl, top_h);
        if (tex == -2) panic("Unknown texture " + tex_name);
        if (bottom.sx != top.sx || bottom.ex != top.ex) panic("Wall start not vertical");
        if ((int16_t)bottom.sx == (int16_t)bottom.ex || (int16_t)top.sx == (int16_t)top.ex) return;
        for (const ScreenLine *l : {&bottom, &top}) {
            if (l->sx < 0 || l->sx >= W) panic("Invalid line start x: " + std::to_string(l->sx));
            if (l->ex < 0 || l->ex >= W) panic("Invalid line end x: " + std::to_string(l->ex));
        }
        const float bottom_delta = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
        const float top_delta = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);

        PlaneAcc acc{this, {floor_flat, ceil_flat}, {floor_sky, ceil_sky}, {sec.floor, sec.ceil}, sec.light};
        const bool full_height = !lower && !upper && !only_occ;
        const int16_t H16 = (int16_t)H, Hm1 = wrap16(H16 - 1);

        Render r;
        r.state = two_sided_mid ? TWOSIDED : SOLID;
        r.bitmap = tex;
        r.light = sec.light;
        r.line = cl;
        r.start_offset = start_offset;
        r.sx = bottom.sx;
        r.ex = bottom.ex;
        r.bottom_h = bottom_h;
        r.top_h = top_h;
        r.off_x = wrap16(as_i16(sd.xoff) + seg_offset);
        r.off_y = wrap16(as_i16(sd.yoff) + wrap16(offset_y));
        r.ext_bottom = lower || (!two_sided_mid && full_height);
        r.ext_top = upper || (!two_sided_mid && full_height);
        r.draw_ceiling = draw_ceiling;
        r.col0 = (uint32_t)colpool.size();

        for (int16_t x = (int16_t)bottom.sx; x < wrap16((int16_t)bottom.ex + 1); x++) {
            if (!hor_ocl[x]) {
                const int16_t bottom_y = as_i16((float)bottom.sy + ((float)x - (float)bottom.sx) * bottom_delta);
                const int16_t top_y = as_i16((float)top.sy + ((float)x - (float)top.sx) * top_delta);
                const int16_t fvo = floor_ocl[x], cvo = ceil_ocl[x];
                int16_t cb = std::min(fvo, bottom_y), ct = std::max(cvo, top_y);
                cb = std::min(Hm1, cb);
                ct = std::max<int16_t>(0, ct);
                const bool in_area = cb >= ct;
                if (in_area) colpool.push_back(drr_col{x, ct, cb, bottom_y, top_y}); // add_column; drawn below if this part is a wall

                if (!two_sided_mid && in_area && (full_height || only_occ)) {
                    bool added = false;
                    if (cb < fvo && cb != Hm1) {
                        acc.point(0, x, cb, fvo);
                        added = true;
                    }
                    if (draw_ceiling && ct > cvo && ct != -1) {
                        acc.point(1, x, cvo, ct);
                        added = true;
                    }
                    if (!added) acc.flush();
                } else if (!two_sided_mid && !in_area && (full_height || only_occ) && fvo > cvo) {
                    if (bottom_y <= cvo) {
                        acc.point(0, x, cvo, fvo);
                        occlude(x);
                    }
                    if (draw_ceiling && top_y >= fvo) {
                        acc.point(1, x, cvo, fvo);
                        occlude(x);
                    }
                }
                if (!two_sided_mid && in_area && only_occ) {
                    floor_ocl[x] = cb;
                    if (draw_ceiling) ceil_ocl[x] = ct;
                }
                if (!two_sided_mid && in_area && lower) floor_ocl[x] = ct;
                if (!two_sided_mid && in_area && upper) ceil_ocl[x] = cb;
            } else {
                acc.flush();
            }
            if (!two_sided_mid && full_height) occlude(x);
        }
        acc.flush();
        r.ncol = (uint32_t)colpool.size() - r.col0;
        // segs.rs:231-258: walls are drawn immediately, column by column; nothing else draws in between, so emitting the
        // whole batch here keeps the draw order.
        if (!two_sided_mid && !only_occ && (phases & DRR_PHASES_WALLS)) emit(r, colpool.data() + r.col0, DRR_PHASE_WALL);
        renders.push_back(r);
    }

    // segs.rs:353-590
    void seg(const SegH &sg) {
        const LineH &ld = lines[sg.line];
        const int fi = sg.dir ? ld.back : ld.front, bi = sg.dir ? ld.front : ld.back;
        if (fi == -1) return;
        const SideH &fs = sides[fi];
        const SectorH &fsec = sectors[fs.sector];
        const float floor_h = (float)fsec.floor;
        float ceil_h = (float)fsec.ceil;
        bool has_pb = false, has_pt = false;
        float pb = 0.0f, pt = 0.0f;
        if (bi != -1) {
            const SectorH &bsec = sectors[sides[bi].sector];
            if (bsec.floor > fsec.floor) {
                has_pb = true;
                pb = (float)bsec.floor;
            }
            if (bsec.ceil < fsec.ceil) {
                has_pt = true;
                pt = (float)bsec.ceil;
            }
        }
        const bool two_sided = (ld.flags & 4) != 0, top_unpeg = (ld.flags & 8) != 0, bottom_unpeg = (ld.flags & 16) != 0;

        const Seg2 view = {rot(sg.v1 - ppos, -pangle), rot(sg.v2 - ppos, -pangle)};
        Seg2 cl;
        float so;
        if (!clip_fov(view, &cl, &so)) return;
        if (cl.s.x < -0.01f) panic("Clipped line x < -0.01");
        const float ph = pfloor + 41.0f;
        const ScreenLine fl = project(cl, floor_h - ph);
        if (fl.sx > fl.ex) return; // back face

        bool fsky, csky;
        const int ff = pick_flat(fsec.floor_flat, &fsky), cf = pick_flat(fsec.ceil_flat, &csky);
        bool draw_ceiling = true;
        if (bi != -1) { // sky hack, segs.rs:463-477
            const SectorH &bsec = sectors[sides[bi].sector];
            if (fsec.ceil_name_has_sky && bsec.ceil_name_has_sky) {
                has_pt = false;
                ceil_h = fminf((float)bsec.ceil, ceil_h);
                draw_ceiling = false;
            }
        }
        auto part = [&](float bh, float th, int32_t oy, int tex, const std::string &tn, bool oo, bool lo, bool upw, bool mid) {
            sidedef_part(cl, so, fs, sg.offset, fsec, ff, fsky, cf, csky, bh, th, oy, tex, tn, oo, lo, upw, draw_ceiling, mid);
        };
        if (!two_sided) {
            part(floor_h - ph, ceil_h - ph, bottom_unpeg ? as_i32(floor_h - ceil_h) : 0, fs.middle, fs.middle_n, false, false, false, false);
        } else {
            part(floor_h - ph, ceil_h - ph, 0, fs.middle, fs.middle_n, true, false, false, false);
            const float mf = has_pb ? pb : floor_h, mc = has_pt ? pt : ceil_h;
            part(mf - ph, mc - ph, 0, fs.middle, fs.middle_n, false, false, false, true);
            if (has_pb) part(floor_h - ph, pb - ph, bottom_unpeg ? as_i32(ceil_h - pb) : 0, fs.lower, fs.lower_n, false, true, false, false);
            if (has_pt) part(pt - ph, ceil_h - ph, top_unpeg ? 0 : as_i32(pt - ceil_h), fs.upper, fs.upper_n, false, false, true, false);
        }
    }

    void walk(int n) { // mod.rs:69-104 (iterative order identical to the recursion: front subtree, then back subtree)
        if (n < 0) {
            auto ss = ssectors[~n];
            for (int i = 0; i < ss.second; i++) seg(segs[ss.first + i]);
            return;
        }
        const NodeH &nd = nodes[n];
        const V2 a = {nd.x, nd.y};
        const bool is_left = left_of(ppos, {a, a + V2{nd.dx, nd.dy}});
        walk(is_left ? nd.left : nd.right);
        walk(is_left ? nd.right : nd.left);
    }

    static bool behind(const Render &r, V2 v) { // bitmap_render.rs:137-165
        const float mn = fminf(r.line.s.x, r.line.e.x), mx = fmaxf(r.line.s.x, r.line.e.x);
        if (mn > v.x) return true;
        return mx > v.x && !left_of(v, r.line);
    }
    void render_deferred(Render &r, const drr_col *pool) { // BitmapRender::render, bitmap_render.rs:101-135
        if (r.state == SOLID || r.state == DRAWN) return;
        if (phases & DRR_PHASES_MASKED) emit(r, pool + r.col0, DRR_PHASE_MASKED);
        r.state = DRAWN;
    }

    void map_objects() { // renderer/map_objects.rs:19-241
        mo_renders.clear();
        mo_cols.clear();
        const int16_t H16 = (int16_t)H, Hm1 = wrap16(H16 - 1);
        for (const ObjH &mo : objects) {
            if (mo.is_null) continue;
            float ang = pangle - mo.angle - PI_F;
            ang += PI_F / 16.0f;
            ang = fmodf(ang, 2.0f * PI_F);
            if (ang < 0.0f) ang += 2.0f * PI_F;
            ang = fmodf(ang, 2.0f * PI_F);
            const uint8_t rotation = as_u8(ang * 8.0f / (2.0f * PI_F));
            auto sp = sprites.find(mo.sprite);
            if (sp == sprites.end() || !sp->second.count(mo.frame)) panic(std::string("Unknown frame for sprite ") + DRR_SPRITE_NAMES[mo.sprite]);
            if (rotation > 7) panic("Invalid rotation");
            const SpriteFrame &sf = sp->second[mo.frame];
            const HostPicture &pic = sf.rotate ? sf.pics[rotation] : sf.pics[0];
            const HostBitmap &bm = bitmaps[pic.bitmap];

            const V2 vpv = rot(mo.pos - ppos, -pangle);
            const Seg2 line = {vpv - V2{0.0f, (float)wrap16(-bm.w) / 2.0f}, vpv - V2{0.0f, (float)bm.w / 2.0f}};
            Seg2 cl;
            float so;
            if (!clip_fov(line, &cl, &so)) continue;
            if (cl.s.x < -0.01f) panic("Clipped line x < -0.01 (map object)");
            const int sector = sector_at(mo.pos);
            if (sector < 0) continue; // "Thing is outside map"
            const int16_t light = mo.full_bright ? (int16_t)255 : sectors[sector].light;
            const float ph = pfloor + 41.0f;
            const int16_t z = sectors[sector].floor;
            float bh = (float)z - ph, th = (float)z + (float)bm.h - 1.0f - ph;
            bh += (float)pic.top_offset - (float)bm.h;
            th += (float)pic.top_offset - (float)bm.h;
            const ScreenLine bottom = project(cl, bh), top = project(cl, th);

            std::fill(top_clip.begin(), top_clip.end(), (int16_t)-1);
            std::fill(bottom_clip.begin(), bottom_clip.end(), H16);
            for (const Render &sg : renders) { // already reversed: back to front (mod.rs:124)
                if (behind(sg, vpv)) continue;
                const drr_col *c = colpool.data() + sg.col0;
                if (sg.state == SOLID) {
                    for (uint32_t i = 0; i < sg.ncol; i++) {
                        if (sg.ext_bottom) bottom_clip[c[i].x] = std::min(bottom_clip[c[i].x], c[i].clipped_top_y);
                        if (sg.ext_top) top_clip[c[i].x] = std::max(top_clip[c[i].x], c[i].clipped_bottom_y);
                    }
                } else if (sg.state == TWOSIDED) {
                    for (uint32_t i = 0; i < sg.ncol; i++) {
                        if (sg.draw_ceiling) top_clip[c[i].x] = std::max(top_clip[c[i].x], c[i].top_y);
                        bottom_clip[c[i].x] = std::min(bottom_clip[c[i].x], c[i].bottom_y);
                    }
                }
            }
            Render r;
            r.state = MAPOBJ;
            r.bitmap = pic.bitmap;
            r.light = light;
            r.line = cl;
            r.start_offset = so;
            r.sx = bottom.sx;
            r.ex = bottom.ex;
            r.bottom_h = bh;
            r.top_h = th;
            r.off_x = r.off_y = 0;
            r.ext_bottom = r.ext_top = r.draw_ceiling = false;
            r.col0 = (uint32_t)mo_cols.size();
            const float bd = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
            const float td = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);
            for (int16_t x = (int16_t)bottom.sx; x < (int16_t)bottom.ex; x++) { // exclusive end (quirk Q6)
                const int16_t by = as_i16((float)bottom.sy + ((float)x - (float)bottom.sx) * bd);
                const int16_t ty = as_i16((float)top.sy + ((float)x - (float)top.sx) * td);
                if (x < 0 || x >= W) panic("map object column outside the screen");
                int16_t ct = std::max(ty, top_clip[x]), cb = std::min(by, bottom_clip[x]);
                ct = std::max<int16_t>(0, ct);
                cb = std::min(Hm1, cb);
                mo_cols.push_back(drr_col{x, ct, cb, by, ty});
            }
            r.ncol = (uint32_t)mo_cols.size() - r.col0;
            mo_renders.push_back(r);
        }
        // stable sort on `start.x as i16`, then reverse (map_objects.rs:216-217, bitmap_render.rs:168-174)
        std::stable_sort(mo_renders.begin(), mo_renders.end(), [](const Render &a, const Render &b) { return as_i16(a.line.s.x) < as_i16(b.line.s.x); });
        std::reverse(mo_renders.begin(), mo_renders.end());
        for (Render &mor : mo_renders) {
            const V2 v = {(mor.line.s.x + mor.line.e.x) / 2.0f, (mor.line.s.y + mor.line.e.y) / 2.0f};
            for (Render &sg : renders)
                if (behind(sg, v)) render_deferred(sg, colpool.data());
            render_deferred(mor, mo_cols.data());
        }
    }

    void emit_view(drr_ctx *c, int view_idx, float x, float y, float angle, float ts, int ph) {
        ctx = c;
        ppos = {x, y};
        pangle = angle;
        timestamp = ts;
        phases = ph;
        pfloor = 0.0f; // game.rs:144-150, 376-389
        const int s = sector_at(ppos);
        if (s >= 0) pfloor = (float)sectors[s].floor;

        renders.clear();
        colpool.clear();
        planes.clear();
        plane_rows.clear();
        hor_ocl.assign(W, 0);
        floor_ocl.assign(W, (int16_t)H); // segs.rs:97-99
        ceil_ocl.assign(W, -1);
        top_clip.resize(W);
        bottom_clip.resize(W);

        drr_view v{x, y, pfloor, angle, cosf(angle), sinf(angle)};
        chk(sink_frame_begin(view_idx, &v));
        try {
            walk((int)nodes.size() - 1);         // A: mod.rs:119-120
            if (phases & DRR_PHASES_PLANES) {    // B: mod.rs:106-116, in creation order
                for (const PlaneH &p : planes) {
                    drr_visplane_hdr h{(int16_t)(p.sky ? DRR_FLAT_SKY : p.flat), p.height, p.light, p.left, p.right, 0};
                    chk(sink_visplane(&h, plane_rows.data() + p.rows + p.left, plane_rows.data() + p.rows + W + p.left));
                }
            }
            std::reverse(renders.begin(), renders.end()); // mod.rs:124
            map_objects();                                  // C
            for (Render &r : renders) render_deferred(r, colpool.data()); // D: segs.rs:593-597
        } catch (...) {
            sink_frame_abort(); // the reference would have panicked: nothing is recorded for this view
            throw;
        }
        chk(sink_frame_end());
    }
};

// ---- C entry points ------------------------------------------------------------------------------------------------
static thread_local std::string g_scene_err;

extern "C" {

const char *drr_scene_last_error(const drr_scene *s) { return s ? s->err.c_str() : g_scene_err.c_str(); }

int drr_scene_load(const char *wad_path, const char *map_name, int width, int height, drr_scene **out) {
    if (!out || !wad_path || !map_name) return DRR_E_INVALID;
    *out = nullptr;
    drr_scene *s = new drr_scene();
    try {
        s->load(wad_path, map_name, width, height);
    } catch (const Panic &p) {
        g_scene_err = p.msg;
        delete s;
        return DRR_E_IO;
    } catch (const std::exception &e) {
        g_scene_err = e.what();
        delete s;
        return DRR_E_NOMEM;
    }
    *out = s;
    return DRR_OK;
}
void drr_scene_free(drr_scene *s) { delete s; }

int drr_scene_upload_assets(drr_scene *s, drr_ctx *ctx) {
    if (!s || !ctx) return DRR_E_INVALID;
    int rc = drr_upload_palette(ctx, s->palette);
    for (size_t i = 0; rc == DRR_OK && i < s->bitmaps.size(); i++) {
        const HostBitmap &b = s->bitmaps[i];
        if (b.w <= 0 || b.h <= 0) continue; // unusable in the reference as well (division by zero)
        rc = drr_upload_bitmap(ctx, (int)i, b.w, b.h, b.px.data());
    }
    for (size_t i = 0; rc == DRR_OK && i < s->flats.size(); i++) rc = drr_upload_flat(ctx, (int)i, s->flats[i].data());
    if (rc == DRR_OK) rc = drr_set_sky(ctx, s->sky_bitmap);
    if (rc != DRR_OK) s->err = std::string("upload_assets: ") + drr_last_error(ctx);
    return rc;
}

int drr_scene_player_start(drr_scene *s, float out_xya[3]) {
    if (!s || !out_xya) return DRR_E_INVALID;
    if (!s->have_start) {
        s->err = "Could not find thing of type 1";
        return DRR_E_PANIC;
    }
    memcpy(out_xya, s->start, sizeof s->start);
    return DRR_OK;
}

int drr_scene_set_tic(drr_scene *s, uint32_t tic, uint64_t seed) {
    if (!s) return DRR_E_INVALID;
    try {
        s->set_tic(tic, seed);
    } catch (const std::exception &e) {
        s->err = e.what();
        return DRR_E_NOMEM;
    }
    return DRR_OK;
}
int drr_scene_world_state(drr_scene *s, int16_t *sector_lights, int32_t *object_states4) {
    if (!s) return DRR_E_INVALID;
    for (size_t i = 0; sector_lights && i < s->sectors.size(); i++) sector_lights[i] = s->sectors[i].light;
    for (size_t i = 0; object_states4 && i < s->objects.size(); i++) {
        const ObjH &o = s->objects[i];
        object_states4[4 * i] = o.sprite;
        object_states4[4 * i + 1] = o.frame;
        object_states4[4 * i + 2] = o.full_bright;
        object_states4[4 * i + 3] = o.is_null;
    }
    return DRR_OK;
}
int drr_scene_counts(drr_scene *s, int *n_sectors, int *n_objects) {
    if (!s) return DRR_E_INVALID;
    if (n_sectors) *n_sectors = (int)s->sectors.size();
    if (n_objects) *n_objects = (int)s->objects.size();
    return DRR_OK;
}

int drr_scene_emit_view(drr_scene *s, drr_ctx *ctx, int view_idx, float x, float y, float angle, float timestamp, int phases) {
    if (!s || !ctx) return DRR_E_INVALID;
    try {
        s->emit_view(ctx, view_idx, x, y, angle, timestamp, phases);
    } catch (const Panic &p) {
        s->err = p.msg;
        return DRR_E_PANIC;
    } catch (const std::exception &e) {
        s->err = e.what();
        return DRR_E_NOMEM;
    }
    return DRR_OK;
}

// Renderer::new(..).render() for n viewpoints at once: the front-end runs on `nthreads` worker threads (each with its own
// copy of the scene's scratch state and its own recorder), the recorded frames are then appended to the context in view
// order.  status[i] (may be NULL) receives DRR_OK or DRR_E_PANIC (the reference would have panicked on that viewpoint:
// nothing is recorded for it).  Returns the first hard error, DRR_OK otherwise.
int drr_scene_emit_views(drr_scene *s, drr_ctx *ctx, int first_view_idx, const float *xya, int n, float timestamp, int phases, int nthreads,
                         int *status) {
    if (!s || !ctx || n < 0 || (n > 0 && !xya)) return DRR_E_INVALID;
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::max(1, std::min(nthreads, std::max(1, n / 8)));
    std::vector<drr_recorder *> recs(nthreads, nullptr);
    std::vector<int> hard(nthreads, DRR_OK);
    std::vector<std::string> msgs(nthreads);
    for (int t = 0; t < nthreads; t++)
        if (drr_recorder_create(ctx, &recs[t]) != DRR_OK) {
            for (auto r : recs) drr_recorder_destroy(r);
            return DRR_E_NOMEM;
        }
    auto work = [&](int t) {
        try {
            drr_scene local(*s); // the scratch state is per thread; maps and assets are only read
            local.rec = recs[t];
            const int lo = (int)((long long)n * t / nthreads), hi = (int)((long long)n * (t + 1) / nthreads);
            for (int i = lo; i < hi; i++) {
                int rc = DRR_OK;
                try {
                    local.emit_view(ctx, first_view_idx + i, xya[3 * i], xya[3 * i + 1], xya[3 * i + 2], timestamp, phases);
                } catch (const Panic &p) {
                    rc = DRR_E_PANIC;
                    if (p.msg.compare(0, 7, "libdrr:") == 0) { // not a reference panic: the library refused something
                        hard[t] = DRR_E_INVALID;
                        msgs[t] = p.msg;
                    }
                }
                if (status) status[i] = rc;
            }
        } catch (const std::exception &e) {
            hard[t] = DRR_E_NOMEM;
            msgs[t] = e.what();
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    int rc = DRR_OK;
    for (int t = 0; t < nthreads && rc == DRR_OK; t++) {
        if (hard[t] != DRR_OK) {
            rc = hard[t];
            s->err = msgs[t];
        } else if ((rc = drr_append(ctx, recs[t])) != DRR_OK) {
            s->err = std::string("drr_append: ") + drr_last_error(ctx);
        }
    }
    for (auto r : recs) drr_recorder_destroy(r);
    return rc;
}

// The map as the flat tables of drr_fe_upload_map, flats resolved for `timestamp` (flats.rs:103-111), then the whole batch
// on the device front-end (csrc/drr_frontend.cuh).
int drr_scene_emit_views_device(drr_scene *s, drr_ctx *ctx, int first_view_idx, const float *xya, int n, float timestamp, int phases, int *status) {
    if (!s || !ctx || n < 0 || (n > 0 && !xya)) return DRR_E_INVALID;
    const uint64_t anim = as_usize(timestamp * 3.0f); // flats.rs:103-111: all that the tables take from the timestamp
    const bool want_things = (phases & DRR_PHASES_MASKED) != 0;
    if (s->fe_ctx == ctx && s->fe_map_id != 0 && s->fe_map_id == drr_fe_map_id(ctx) && s->fe_anim == anim && s->fe_things == want_things && s->fe_world == s->world_version) {
        const int rc = drr_fe_emit_views(ctx, first_view_idx, xya, n, phases, status); // the context still has this map
        if (rc != DRR_OK) s->err = std::string("device front-end: ") + drr_last_error(ctx);
        return rc;
    }
    std::vector<drr_fe_node> nodes(s->nodes.size());
    for (size_t i = 0; i < nodes.size(); i++) nodes[i] = drr_fe_node{s->nodes[i].x, s->nodes[i].y, s->nodes[i].dx, s->nodes[i].dy, s->nodes[i].right, s->nodes[i].left};
    std::vector<drr_fe_subsector> ss(s->ssectors.size());
    for (size_t i = 0; i < ss.size(); i++) ss[i] = drr_fe_subsector{s->ssectors[i].first, s->ssectors[i].second};
    std::vector<drr_fe_seg> segs(s->segs.size());
    for (size_t i = 0; i < segs.size(); i++) {
        const SegH &g = s->segs[i];
        segs[i] = drr_fe_seg{g.v1.x, g.v1.y, g.v2.x, g.v2.y, g.line, (int16_t)(g.dir ? 1 : 0), g.offset};
    }
    std::vector<drr_fe_linedef> lines(s->lines.size());
    for (size_t i = 0; i < lines.size(); i++) lines[i] = drr_fe_linedef{s->lines[i].front, s->lines[i].back, (int32_t)s->lines[i].flags};
    std::vector<drr_fe_sidedef> sides(s->sides.size());
    for (size_t i = 0; i < sides.size(); i++) {
        const SideH &d = s->sides[i];
        sides[i] = drr_fe_sidedef{d.xoff, d.yoff, d.upper, d.lower, d.middle, d.sector}; // (zero-sized bitmaps were never uploaded: the library reports them)
    }
    std::vector<drr_fe_sector> sectors(s->sectors.size());
    s->timestamp = timestamp;
    for (size_t i = 0; i < sectors.size(); i++) {
        const SectorH &c = s->sectors[i];
        drr_fe_sector o{};
        o.floor_height = c.floor;
        o.ceiling_height = c.ceil;
        o.light_level = c.light;
        o.ceiling_name_has_sky = c.ceil_name_has_sky ? 1 : 0;
        auto pick = [&](const FlatRef &fr, int16_t *id, int16_t *sky) { // pick_flat without the panic: -2 marks a missing lump
            const size_t k = fr.ids.size() == 1 ? 0 : (size_t)(as_usize(timestamp * 3.0f) % fr.ids.size());
            *id = (int16_t)(fr.ids[k] < 0 ? -2 : fr.ids[k]);
            *sky = fr.is_sky[k] ? 1 : 0;
        };
        pick(c.floor_flat, &o.floor_flat, &o.floor_is_sky);
        pick(c.ceil_flat, &o.ceiling_flat, &o.ceiling_is_sky);
        sectors[i] = o;
    }
    // map objects (map_objects.rs:25-50): sprite frame and sector do not depend on the viewpoint
    std::vector<drr_fe_thing> things;
    bool every_view_panics = false;
    if (phases & DRR_PHASES_MASKED)
        for (const ObjH &mo : s->objects) {
            if (mo.is_null) continue;
            auto sp = s->sprites.find(mo.sprite);
            if (sp == s->sprites.end() || !sp->second.count(mo.frame)) { // "Unknown frame for sprite": panics before anything view-dependent
                every_view_panics = true;
                break;
            }
            const SpriteFrame &sf = sp->second[mo.frame];
            drr_fe_thing t{};
            t.x = mo.pos.x;
            t.y = mo.pos.y;
            t.angle = mo.angle;
            t.sector = s->sector_at(mo.pos);
            t.full_bright = mo.full_bright ? 1 : 0;
            t.rotate = sf.rotate ? 1 : 0;
            for (int r = 0; r < (sf.rotate ? 8 : 1); r++) {
                t.bitmap[r] = sf.pics[r].bitmap;
                t.top_offset[r] = sf.pics[r].top_offset;
            }
            things.push_back(t);
        }
    if (every_view_panics) {
        for (int i = 0; status && i < n; i++) status[i] = DRR_E_PANIC;
        return DRR_OK;
    }
    drr_fe_map m{};
    m.nodes = nodes.data();
    m.n_nodes = (int32_t)nodes.size();
    m.subsectors = ss.data();
    m.n_subsectors = (int32_t)ss.size();
    m.segs = segs.data();
    m.n_segs = (int32_t)segs.size();
    m.linedefs = lines.data();
    m.n_linedefs = (int32_t)lines.size();
    m.sidedefs = sides.data();
    m.n_sidedefs = (int32_t)sides.size();
    m.sectors = sectors.data();
    m.n_sectors = (int32_t)sectors.size();
    m.things = things.data();
    m.n_things = (int32_t)things.size();
    int rc = drr_fe_upload_map(ctx, &m);
    s->fe_map_id = 0;
    if (rc == DRR_OK) {
        s->fe_ctx = ctx;
        s->fe_map_id = drr_fe_map_id(ctx);
        s->fe_anim = anim;
        s->fe_world = s->world_version;
        s->fe_things = want_things;
        rc = drr_fe_emit_views(ctx, first_view_idx, xya, n, phases, status);
    }
    if (rc != DRR_OK) s->err = std::string("device front-end: ") + drr_last_error(ctx);
    return rc;
}

} // extern "C"
