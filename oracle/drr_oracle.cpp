// drr_oracle.cpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A literal, single-threaded C++ restatement of freewilll/doom-rust-renderer's renderer
// (reference paths are relative to the reference repo root). It exists so that the CUDA
// path can be checked bit-for-bit; it is NOT part of the product. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// PARITY STATUS: "parity unpinned" by the reference itself -- the reference has no tests,
// golden vectors or fixtures, and it cannot be compiled here (no
// rustc/cargo, no SDL2). Faithfulness is argued by the line-by-line citations below and by
// the closure tests in tests/ (direct render == replay of the recorded leaf calls == GPU).
//
// Arithmetic contract: IEEE f32, no FMA contraction (build with
// -ffp-contract=off, no fast-math), Rust `as` casts (truncate, saturate, NaN->0), i16
// arithmetic wraps (release build), sinf/cosf/sqrtf from the host libm.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -shared -fPIC)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../data/drr_info_table.inc"

namespace orc {

// ---------------------------------------------------------------------------------------
// Rust scalar semantics
// ---------------------------------------------------------------------------------------
static inline int16_t f2i16(float f) {  // `f as i16`
    if (f != f) return 0;
    if (f <= -32768.0f) return INT16_MIN;
    if (f >= 32767.0f) return INT16_MAX;
    return (int16_t)f;
}
static inline int32_t f2i32(float f) {  // `f as i32`
    if (f != f) return 0;
    if (f <= -2147483648.0f) return INT32_MIN;
    if (f >= 2147483648.0f) return INT32_MAX;
    return (int32_t)f;
}
static inline uint8_t f2u8(float f) {  // `f as u8`
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
static inline uint64_t f2usize(float f) {  // `f as usize`
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)f;
}
static inline int16_t w16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }  // wrapping i16
static inline int16_t add16(int16_t a, int16_t b) { return w16((int32_t)a + (int32_t)b); }
static inline int16_t sub16(int16_t a, int16_t b) { return w16((int32_t)a - (int32_t)b); }
static inline int16_t mul16(int16_t a, int16_t b) { return w16((int32_t)a * (int32_t)b); }
static inline int32_t sub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline float fmin_rs(float a, float b) { return fminf(a, b); }  // f32::min ignores NaN
static inline float fmax_rs(float a, float b) { return fmaxf(a, b); }

static const float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI

struct Error {
    std::string msg;
};
[[noreturn]] static void rs_panic(const std::string &m) { throw Error{m}; }

// ---------------------------------------------------------------------------------------
// map/vertexes.rs, geometry.rs
// ---------------------------------------------------------------------------------------
struct Vertex {
    float x, y;
};
struct Line {
    Vertex start, end;
};
static inline Vertex vsub(const Vertex &a, const Vertex &b) { return {a.x - b.x, a.y - b.y}; }   // vertexes.rs:58-67
static inline Vertex vadd(const Vertex &a, const Vertex &b) { return {a.x + b.x, a.y + b.y}; }   // vertexes.rs:47-56
static inline Vertex rotate(const Vertex &v, float angle) {                                      // vertexes.rs:20-25
    Vertex r;
    r.x = v.x * cosf(angle) - v.y * sinf(angle);
    r.y = v.y * cosf(angle) + v.x * sinf(angle);
    return r;
}
static inline float cross_product(const Vertex &a, const Vertex &b) { return a.x * b.y - a.y * b.x; }  // :27-29
static inline bool is_left_of_line(const Vertex &v, const Line &l) {                                   // :32-34
    return cross_product(vsub(v, l.start), vsub(l.end, l.start)) <= 0.0f;
}
static inline float distance_to(const Vertex &a, const Vertex &b) {  // :36-38
    float dx = a.x - b.x, dy = a.y - b.y;
    return sqrtf(dx * dx + dy * dy);
}
static inline float line_length(const Line &l) {  // geometry.rs:84-86
    float dx = l.start.x - l.end.x, dy = l.start.y - l.end.y;
    return sqrtf(dx * dx + dy * dy);
}
// geometry.rs:56-82. Returns false when "parallel".
static bool intersection(const Line &a, const Line &b, Vertex *out) {
    float x1 = a.start.x, y1 = a.start.y, x2 = a.end.x, y2 = a.end.y;
    float x3 = b.start.x, y3 = b.start.y, x4 = b.end.x, y4 = b.end.y;
    float quot = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (fabsf(quot) < 0.001f) return false;
    float invquot = 1.0f / quot;
    float px = invquot * ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4));
    float py = invquot * ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4));
    *out = {px, py};
    return true;
}

// ---------------------------------------------------------------------------------------
// wad.rs
// ---------------------------------------------------------------------------------------
static std::string upper(const std::string &s) {
    std::string r = s;
    for (auto &c : r)
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return r;
}
struct DirEntry {
    int16_t index;
    std::string name;
    uint32_t offset, size;
};
enum MapLump { Things = 1, Linedefs, Sidedefs, Vertexes, SegsL, Ssectors, Nodes, Sectors, Reject, Blockmap };  // wad.rs:8-19

struct WadFile {
    std::vector<uint8_t> file;
    std::vector<DirEntry> dirs_list;
    std::map<std::string, int> dirs_map;  // name -> index in dirs_list (last duplicate wins, wad.rs:153-155)
    int16_t first_sprite_lump = -1, last_sprite_lump = -1;

    uint8_t u8(size_t o) const {
        if (o >= file.size()) rs_panic("wad read out of bounds");
        return file[o];
    }
    int16_t read_i16(size_t o) const { return (int16_t)(uint16_t)(u8(o) | (u8(o + 1) << 8)); }          // :185-187
    float read_f32_from_i16(size_t o) const { return (float)read_i16(o); }                              // :189-191
    uint32_t read_u32(size_t o) const {                                                                 // :193-195
        return (uint32_t)u8(o) | ((uint32_t)u8(o + 1) << 8) | ((uint32_t)u8(o + 2) << 16) | ((uint32_t)u8(o + 3) << 24);
    }
    std::string read_lump_name(size_t o) const {  // :112-126
        std::string s;
        if (u8(o + 7) == 0) {
            for (int i = 0; i < 8 && u8(o + i) != 0; i++) s.push_back((char)u8(o + i));
        } else {
            for (int i = 0; i < 8; i++) s.push_back((char)u8(o + i));
        }
        return s;
    }
    const DirEntry *get_dir_entry(const std::string &name) const {  // :166-172 (nullptr == Err)
        auto it = dirs_map.find(upper(name));
        if (it == dirs_map.end()) return nullptr;
        return &dirs_list[it->second];
    }
    const DirEntry &get_dir_entry_for_map_lump(const std::string &map_name, int lump) const {  // :175-183
        std::string m = upper(map_name);
        for (size_t i = 0; i < dirs_list.size(); i++) {
            if (dirs_list[i].name == m) {
                if (i + lump >= dirs_list.size()) rs_panic("map lump index out of range");
                return dirs_list[i + lump];
            }
        }
        rs_panic("Could not find map " + map_name);
    }
    void load(std::vector<uint8_t> &&bytes) {  // :86-109, :128-157
        file = std::move(bytes);
        if (file.size() < 12) rs_panic("wad too short");
        if (std::string((const char *)file.data(), 4) != "IWAD") rs_panic("Unhandled WAD file type");
        uint32_t lump_count = read_u32(4), dir_offset = read_u32(8);
        for (uint32_t i = 0; i < lump_count; i++) {
            size_t e = (size_t)dir_offset + (size_t)i * 16;
            DirEntry d;
            d.index = (int16_t)i;
            d.offset = read_u32(e);
            d.size = read_u32(e + 4);
            d.name = upper(read_lump_name(e + 8));
            dirs_map[d.name] = (int)dirs_list.size();
            dirs_list.push_back(d);
        }
        const DirEntry *s = get_dir_entry("S_START"), *e = get_dir_entry("S_END");
        if (!s || !e) rs_panic("S_START/S_END missing");
        first_sprite_lump = s->index;
        last_sprite_lump = e->index;
    }
};

// ---------------------------------------------------------------------------------------
// graphics/*
// ---------------------------------------------------------------------------------------
struct Bitmap {  // graphics/bitmap.rs:11-15 ; texel -1 == None
    int id = -1;
    int16_t width = 0, height = 0;
    std::vector<int16_t> pixels;  // row-major [y*width + x]
    int16_t &at(int y, int x) { return pixels[(size_t)y * (size_t)width + (size_t)x]; }
    int16_t get(int y, int x) const {
        if (y < 0 || y >= height || x < 0 || x >= width) rs_panic("bitmap index out of bounds");
        return pixels[(size_t)y * (size_t)width + (size_t)x];
    }
};
struct Picture {  // graphics/pictures.rs:20-26
    std::string name;
    std::shared_ptr<Bitmap> bitmap;
    int16_t left_offset = 0, top_offset = 0;
};
struct Flat {  // graphics/flats.rs:19-22
    int id = -1;
    std::string name;
    uint8_t pixels[64][64];
};
struct Color {
    uint8_t r, g, b;
};

struct Assets {
    const WadFile *wad = nullptr;
    Color palette[256];                                        // palette.rs:11-28
    std::vector<std::shared_ptr<Bitmap>> bitmaps;              // registry (id == index)
    std::vector<std::shared_ptr<Flat>> flat_list;              // registry (id == index)
    std::map<std::string, std::shared_ptr<Flat>> flat_map;     // flats.rs:13
    std::map<std::string, std::vector<std::string>> animated;  // flats.rs:15,30-89
    std::map<std::string, std::shared_ptr<Picture>> pictures;  // pictures.rs:14

    struct Patch {
        int16_t origin_x, origin_y, patch_number;
    };
    struct TexDef {
        int16_t width, height;
        std::vector<Patch> patches;
        std::shared_ptr<Bitmap> bitmap;  // loaded texture (lazy)
    };
    std::map<std::string, TexDef> texdefs;  // textures.rs:51
    std::vector<std::string> pnames;        // textures.rs:53

    struct SpriteFrame {
        bool rotate = false;
        std::vector<std::shared_ptr<Picture>> pictures;
    };
    std::map<int, std::map<int, SpriteFrame>> sprites;  // sprite index -> frame -> SpriteFrame

    std::shared_ptr<Bitmap> register_bitmap(std::shared_ptr<Bitmap> b) {
        b->id = (int)bitmaps.size();
        bitmaps.push_back(b);
        return b;
    }

    void load_palette() {
        const DirEntry *d = wad->get_dir_entry("PLAYPAL");
        if (!d) rs_panic("PLAYPAL missing");
        for (int i = 0; i < 256; i++)
            palette[i] = {wad->u8(d->offset + i * 3), wad->u8(d->offset + i * 3 + 1), wad->u8(d->offset + i * 3 + 2)};
    }

    // pictures.rs:66-96,100-126
    std::shared_ptr<Picture> picture_new(const std::string &name) {
        const DirEntry *d = wad->get_dir_entry(name);
        if (!d) return nullptr;
        size_t off = d->offset;
        auto bm = std::make_shared<Bitmap>();
        bm->width = wad->read_i16(off);
        bm->height = wad->read_i16(off + 2);
        if (bm->width < 0 || bm->height < 0) rs_panic("negative picture size");
        bm->pixels.assign((size_t)bm->width * (size_t)bm->height, -1);
        auto pic = std::make_shared<Picture>();
        pic->name = name;
        pic->left_offset = wad->read_i16(off + 4);
        pic->top_offset = wad->read_i16(off + 6);
        for (int column = 0; column < bm->width; column++) {
            size_t co = off + wad->read_u32(off + (size_t)column * 4 + 8);
            for (;;) {
                uint8_t y_offset = wad->u8(co);
                if (y_offset == 0xff) break;
                uint8_t length = wad->u8(co + 1);
                for (int row = 0; row < length; row++) {
                    uint8_t value = wad->u8(co + row + 3);
                    int y = row + y_offset;
                    if (y >= bm->height) rs_panic("picture post out of bounds");
                    bm->at(y, column) = value;
                }
                co += (size_t)length + 4;
            }
        }
        pic->bitmap = register_bitmap(bm);
        return pic;
    }
    std::shared_ptr<Picture> picture_get(const std::string &name) {  // pictures.rs:38-47
        auto it = pictures.find(name);
        if (it != pictures.end()) return it->second;
        auto p = picture_new(name);
        if (!p) return nullptr;
        pictures[name] = p;
        return p;
    }
    std::shared_ptr<Picture> picture_mirror(const Picture &p) {  // pictures.rs:129-147
        auto bm = std::make_shared<Bitmap>(*p.bitmap);
        for (int y = 0; y < bm->height; y++)
            for (int x = 0; x < bm->width / 2; x++) std::swap(bm->at(y, x), bm->at(y, bm->width - 1 - x));
        auto q = std::make_shared<Picture>(p);
        q->bitmap = register_bitmap(bm);
        return q;
    }

    // flats.rs:14-111
    void init_flats() {
        static const std::vector<std::vector<std::string>> lists = {
            {"NUKAGE1", "NUKAGE2", "NUKAGE3"},
            {"FWATER1", "FWATER2", "FWATER3", "FWATER4"},
            {"SWATER1", "SWATER2", "SWATER3", "SWATER4"},
            {"LAVA1", "LAVA2", "LAVA3", "LAVA4"},
            {"BLOOD1", "BLOOD2", "BLOOD3"},
            {"RROCK05", "RROCK06", "RROCK07", "RROCK08"},
            {"SLIME01", "SLIME02", "SLIME03", "SLIME04"},
            {"SLIME05", "SLIME06", "SLIME07", "SLIME08"},
            {"SLIME09", "SLIME10", "SLIME11", "SLIME12"},
        };
        for (auto &l : lists)
            for (auto &n : l) animated[n] = l;
    }
    std::shared_ptr<Flat> flat_get(const std::string &name) {  // flats.rs:92-100,116-137
        auto it = flat_map.find(name);
        if (it != flat_map.end()) return it->second;
        const DirEntry *d = wad->get_dir_entry(name);
        if (!d) rs_panic("Could not find flat " + name);
        auto f = std::make_shared<Flat>();
        f->name = name;
        for (int y = 0; y < 64; y++)
            for (int x = 0; x < 64; x++) f->pixels[y][x] = wad->u8(d->offset + y * 64 + x);
        f->id = (int)flat_list.size();
        flat_list.push_back(f);
        flat_map[name] = f;
        return f;
    }
    std::shared_ptr<Flat> flat_get_animated(const std::string &name, float timestamp) {  // flats.rs:103-111
        auto it = animated.find(name);
        if (it != animated.end()) {
            uint64_t cycle = f2usize(timestamp * 3.0f) % it->second.size();
            return flat_get(it->second[cycle]);
        }
        return flat_get(name);
    }

    // textures.rs:131-255
    void load_texture_list(const DirEntry &d) {
        size_t base = d.offset;
        uint32_t count = wad->read_u32(base);
        for (uint32_t i = 0; i < count; i++) {
            size_t off = base + wad->read_u32(base + 4 + 4 * (size_t)i);
            std::string name = wad->read_lump_name(off);
            TexDef td;
            td.width = wad->read_i16(off + 12);
            td.height = wad->read_i16(off + 14);
            int16_t patch_count = wad->read_i16(off + 20);
            for (int j = 0; j < patch_count; j++) {
                size_t po = off + 22 + (size_t)j * 10;
                td.patches.push_back({wad->read_i16(po), wad->read_i16(po + 2), wad->read_i16(po + 4)});
            }
            texdefs[upper(name)] = td;
        }
    }
    void init_textures() {
        const DirEntry *pn = wad->get_dir_entry("PNAMES");
        if (!pn) rs_panic("PNAMES missing");
        uint32_t count = wad->read_u32(pn->offset);
        for (uint32_t i = 0; i < count; i++) pnames.push_back(wad->read_lump_name(pn->offset + 4 + (size_t)i * 8));
        const DirEntry *t1 = wad->get_dir_entry("TEXTURE1");
        if (!t1) rs_panic("TEXTURE1 missing");
        load_texture_list(*t1);
        if (const DirEntry *t2 = wad->get_dir_entry("TEXTURE2")) load_texture_list(*t2);
    }
    std::shared_ptr<Bitmap> texture_get(const std::string &name) {  // textures.rs:154-179, load :74-103
        auto it = texdefs.find(upper(name));
        if (it == texdefs.end()) rs_panic("Unknown texture " + name);
        TexDef &td = it->second;
        if (td.bitmap) return td.bitmap;
        auto bm = std::make_shared<Bitmap>();
        bm->width = td.width;
        bm->height = td.height;
        if (bm->width < 0 || bm->height < 0) rs_panic("negative texture size");
        bm->pixels.assign((size_t)bm->width * (size_t)bm->height, -1);
        for (auto &patch : td.patches) {
            if (patch.patch_number < 0 || (size_t)patch.patch_number >= pnames.size()) rs_panic("bad patch number");
            // Patch::get_picture builds a fresh Picture per patch (textures.rs:58-68), not via the Pictures cache.
            auto pic = picture_new(pnames[patch.patch_number]);
            if (!pic) rs_panic("missing patch " + pnames[patch.patch_number]);
            const Bitmap &pb = *pic->bitmap;
            for (int x = 0; x < pb.width; x++) {
                for (int y = 0; y < pb.height; y++) {
                    int16_t value = pb.pixels[(size_t)y * pb.width + x];
                    int16_t px = add16((int16_t)x, patch.origin_x), py = add16((int16_t)y, patch.origin_y);
                    if (px >= 0 && px < bm->width && py >= 0 && py < bm->height)
                        bm->at(py, px) = value;  // unconditional: None punches holes (quirk Q1)
                }
            }
        }
        td.bitmap = register_bitmap(bm);
        return td.bitmap;
    }

    // sprites.rs:26-97
    void init_sprites() {
        for (int sid = 0; sid < 138; sid++) {
            std::string sprite_name = DRR_SPRITE_NAMES[sid];
            std::map<int, std::map<int, std::shared_ptr<Picture>>> found;  // frame -> rotation -> picture
            for (int index = wad->first_sprite_lump; index < wad->last_sprite_lump; index++) {
                const DirEntry &d = wad->dirs_list[index];
                if (d.name.compare(0, sprite_name.size(), sprite_name) != 0) continue;
                if (d.name.size() < 6) rs_panic("sprite lump name too short: " + d.name);
                auto picture = picture_get(d.name);
                if (!picture) rs_panic("sprite picture missing");
                uint8_t frame = (uint8_t)(d.name[4] - 65), rotation = (uint8_t)(d.name[5] - 48);
                found[frame][rotation] = picture;
                if (d.name.size() > 6) {
                    if (d.name.size() < 8) rs_panic("sprite lump name has 7 characters: " + d.name);
                    uint8_t frame2 = (uint8_t)(d.name[6] - 65), rotation2 = (uint8_t)(d.name[7] - 48);
                    found[frame2][rotation2] = picture_mirror(*picture);
                }
            }
            auto &sprite = sprites[sid];
            for (auto &fr : found) {
                SpriteFrame sf;
                sf.rotate = fr.second.size() != 1;
                if (sf.rotate) {
                    if (fr.second.size() != 8) rs_panic("Got something other than 8 rotations for " + sprite_name);
                    for (int rot = 1; rot < 9; rot++) {
                        auto it = fr.second.find(rot);
                        if (it == fr.second.end()) rs_panic("missing rotation");
                        sf.pictures.push_back(it->second);
                    }
                } else {
                    auto it = fr.second.find(0);
                    if (it == fr.second.end()) rs_panic("single rotation is not 0");
                    sf.pictures.push_back(it->second);
                }
                sprite[fr.first] = sf;
            }
        }
    }
    std::shared_ptr<Picture> sprite_get_picture(int sprite_id, uint8_t frame_id, uint8_t rotation) {  // sprites.rs:99-117
        auto &sprite = sprites[sprite_id];
        auto it = sprite.find(frame_id);
        if (it == sprite.end()) rs_panic("Unknown frame for sprite " + std::string(DRR_SPRITE_NAMES[sprite_id]));
        if (rotation > 7) rs_panic("Invalid rotation");
        return it->second.rotate ? it->second.pictures[rotation] : it->second.pictures[0];
    }
};

// ---------------------------------------------------------------------------------------
// map/*
// ---------------------------------------------------------------------------------------
struct Sector {
    int16_t floor_height, ceiling_height;
    std::string floor_texture, ceiling_texture;
    int16_t light_level, special_type, tag_number;
};
struct Sidedef {
    float x_offset, y_offset;
    std::string upper_texture, lower_texture, middle_texture;
    int sector;
};
struct Linedef {
    int start_vertex, end_vertex;
    int16_t flags;
    int front_sidedef, back_sidedef;  // -1 == None
};
struct Seg {
    int start_vertex, end_vertex, linedef;
    bool direction;
    int16_t offset;
};
struct SubSector {
    std::vector<int> segs;
};
struct Node {
    float x, y, dx, dy;
    bool right_is_subsector, left_is_subsector;
    int right_child, left_child;
};
struct Thing {
    float x, y, angle;
    int16_t thing_type, flags;
};
struct MapObject {  // map_objects.rs:11-17; sprite / frame / full_bright / is_null mirror STATES[state]
    int sprite;
    uint8_t frame;
    bool full_bright, is_null;
    Vertex position;
    float angle;
    int state = 0;  // StateId of the current state (spawn state at tic 0)
};

struct Map {
    std::vector<Thing> things;
    std::vector<Vertex> vertexes;
    std::vector<Sector> sectors;
    std::vector<Sidedef> sidedefs;
    std::vector<Linedef> linedefs;
    std::vector<Seg> segs;
    std::vector<SubSector> subsectors;
    std::vector<Node> nodes;
    std::vector<MapObject> objects;

    template <class T>
    static const T &idx(const std::vector<T> &v, long i, const char *what) {
        if (i < 0 || (size_t)i >= v.size()) rs_panic(std::string("index out of bounds: ") + what);
        return v[(size_t)i];
    }

    void load(const WadFile &w, const std::string &name) {  // map/mod.rs:48-78
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Things);  // things.rs:27-46
            for (size_t i = 0; i < d.size / 10; i++) {
                size_t o = d.offset + i * 10;
                Thing t;
                t.x = w.read_f32_from_i16(o);
                t.y = w.read_f32_from_i16(o + 2);
                t.angle = w.read_f32_from_i16(o + 4) * (PI_F / 180.0f);  // f32::to_radians
                t.thing_type = w.read_i16(o + 6);
                t.flags = w.read_i16(o + 8);
                things.push_back(t);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Vertexes);  // vertexes.rs:69-84
            for (size_t i = 0; i < d.size / 4; i++)
                vertexes.push_back({w.read_f32_from_i16(d.offset + i * 4), w.read_f32_from_i16(d.offset + i * 4 + 2)});
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Sectors);  // sectors.rs:19-44
            for (size_t i = 0; i < d.size / 26; i++) {
                size_t o = d.offset + i * 26;
                Sector s;
                s.floor_height = w.read_i16(o);
                s.ceiling_height = w.read_i16(o + 2);
                s.floor_texture = w.read_lump_name(o + 4);
                s.ceiling_texture = w.read_lump_name(o + 12);
                s.light_level = w.read_i16(o + 20);
                s.special_type = w.read_i16(o + 22);
                s.tag_number = w.read_i16(o + 24);
                sectors.push_back(s);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Sidedefs);  // sidedefs.rs:19-44
            for (size_t i = 0; i < d.size / 30; i++) {
                size_t o = d.offset + i * 30;
                Sidedef s;
                s.x_offset = w.read_f32_from_i16(o);
                s.y_offset = w.read_f32_from_i16(o + 2);
                s.upper_texture = w.read_lump_name(o + 4);
                s.lower_texture = w.read_lump_name(o + 12);
                s.middle_texture = w.read_lump_name(o + 20);
                s.sector = (int)(uint16_t)w.read_i16(o + 28);  // `as usize` of an i16 index
                idx(sectors, w.read_i16(o + 28), "sidedef sector");
                sidedefs.push_back(s);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Linedefs);  // linedefs.rs:34-75
            for (size_t i = 0; i < d.size / 14; i++) {
                size_t o = d.offset + i * 14;
                Linedef l;
                l.start_vertex = w.read_i16(o);
                l.end_vertex = w.read_i16(o + 2);
                idx(vertexes, l.start_vertex, "linedef v1");
                idx(vertexes, l.end_vertex, "linedef v2");
                l.flags = w.read_i16(o + 4);
                l.front_sidedef = w.read_i16(o + 10);
                l.back_sidedef = w.read_i16(o + 12);
                if (l.front_sidedef != -1) idx(sidedefs, l.front_sidedef, "front sidedef");
                if (l.back_sidedef != -1) idx(sidedefs, l.back_sidedef, "back sidedef");
                linedefs.push_back(l);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, SegsL);  // map/segs.rs:17-42
            for (size_t i = 0; i < d.size / 12; i++) {
                size_t o = d.offset + i * 12;
                Seg s;
                s.start_vertex = w.read_i16(o);
                s.end_vertex = w.read_i16(o + 2);
                s.linedef = w.read_i16(o + 6);
                idx(vertexes, s.start_vertex, "seg v1");
                idx(vertexes, s.end_vertex, "seg v2");
                idx(linedefs, s.linedef, "seg linedef");
                s.direction = w.read_i16(o + 8) != 0;
                s.offset = w.read_i16(o + 10);
                segs.push_back(s);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Ssectors);  // subsectors.rs:10-33
            for (size_t i = 0; i < d.size / 4; i++) {
                int16_t seg_count = w.read_i16(d.offset + i * 4), first = w.read_i16(d.offset + i * 4 + 2);
                SubSector ss;
                for (int k = first; k < first + seg_count; k++) {
                    idx(segs, k, "subsector seg");
                    ss.segs.push_back(k);
                }
                subsectors.push_back(ss);
            }
        }
        {
            const DirEntry &d = w.get_dir_entry_for_map_lump(name, Nodes);  // nodes.rs:44-83
            for (size_t i = 0; i < d.size / 28; i++) {
                size_t o = d.offset + i * 28;
                Node n;
                n.x = w.read_f32_from_i16(o);
                n.y = w.read_f32_from_i16(o + 2);
                n.dx = w.read_f32_from_i16(o + 4);
                n.dy = w.read_f32_from_i16(o + 6);
                auto child = [&](int16_t index, bool *is_ss, int *out) {  // nodes.rs:17-27
                    *is_ss = (index & INT16_MIN) == INT16_MIN;
                    *out = index & 0x7fff;
                    if (*is_ss)
                        idx(subsectors, *out, "node subsector child");
                    else if ((size_t)*out >= nodes.size())
                        rs_panic("node child index not yet loaded");
                };
                child(w.read_i16(o + 24), &n.right_is_subsector, &n.right_child);
                child(w.read_i16(o + 26), &n.left_is_subsector, &n.left_child);
                nodes.push_back(n);
            }
            if (nodes.empty()) rs_panic("map has no nodes");
        }
        // map_objects.rs:25-50
        for (auto &t : things) {
            if ((t.thing_type >= 1 && t.thing_type <= 4) || t.thing_type == 11) continue;
            const DrrThingInfo *info = nullptr;
            for (auto &ti : DRR_THING_INFOS)
                if (ti.doomednum == t.thing_type) info = &ti;
            if (!info) rs_panic("unknown thing type " + std::to_string(t.thing_type));
            objects.push_back({info->sprite, info->frame, info->full_bright != 0, info->is_null != 0, {t.x, t.y}, t.angle,
                               (int)DRR_THING_SPAWN_STATE[info - DRR_THING_INFOS]});
        }
    }
};

// ---------------------------------------------------------------------------------------
// The time axis: thinkers.rs, lights.rs, map_objects.rs:63-95, game.rs:456-482.
// The reference draws its random numbers from rand::thread_rng(), so it is not reproducible from run to run; here every
// draw comes from ONE PCG32 stream (seed given by the caller), consumed in the reference's own order: thinker construction
// in list order (sector thinkers in sector order, then the map objects), then list order every tic.
// gen_range(lo..hi) = lo + next_u32() % (hi - lo).
// ---------------------------------------------------------------------------------------
struct Pcg32 {
    uint64_t state = 0, inc = 1;
    explicit Pcg32(uint64_t seed) {
        inc = (0xda3e39cb94b95bdbull << 1) | 1u;
        next();
        state += seed;
        next();
    }
    uint32_t next() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    int16_t gen_range(int lo, int hi) { return (int16_t)(lo + (int)(next() % (uint32_t)(hi - lo))); }
};

struct Thinkers {
    enum Kind { LIGHT_FLASH, STROBE_FLASH, GLOWING_LIGHT, FIRE_FLICKER, MAP_OBJECT };
    struct T {
        Kind kind;
        int target;  // sector or map object
        int16_t min_light = 0, max_light = 0, min_time = 0, max_time = 0, dark_time = 0, bright_time = 0, count = 0;
        bool going_up = false;
    };
    std::vector<T> list;

    // lights.rs:14-43
    static int16_t find_min_surrounding_light(const Map &map, int sector_id, int16_t max) {
        int16_t light_level = max;
        for (const Linedef &l : map.linedefs) {
            if (l.front_sidedef != -1 && map.sidedefs[l.front_sidedef].sector == sector_id && l.back_sidedef != -1)
                light_level = std::min(light_level, map.sectors[map.sidedefs[l.back_sidedef].sector].light_level);
            if (l.back_sidedef != -1 && map.sidedefs[l.back_sidedef].sector == sector_id && l.front_sidedef != -1)
                light_level = std::min(light_level, map.sectors[map.sidedefs[l.front_sidedef].sector].light_level);
        }
        return light_level;
    }

    void init(Map &map, Pcg32 &rng) {  // thinkers.rs:14-91
        list.clear();
        for (int i = 0; i < (int)map.sectors.size(); i++) {
            const Sector &sec = map.sectors[i];
            T t{};
            t.target = i;
            auto strobe = [&](int16_t dark_time, bool in_sync) {  // lights.rs:112-141
                t.kind = STROBE_FLASH;
                t.min_light = find_min_surrounding_light(map, i, sec.light_level);
                t.max_light = sec.light_level;
                if (t.min_light == t.max_light) t.min_light = 0;
                t.count = in_sync ? (int16_t)1 : rng.gen_range(1, 9);
                t.dark_time = dark_time;
                t.bright_time = 5;  // STROBE_BRIGHT
                list.push_back(t);
            };
            switch (sec.special_type) {
            case 1:  // lights.rs:58-78
                t.kind = LIGHT_FLASH;
                t.min_light = find_min_surrounding_light(map, i, sec.light_level);
                t.max_light = sec.light_level;
                t.min_time = 7;
                t.max_time = 64;
                t.count = rng.gen_range(1, t.max_time + 1);
                list.push_back(t);
                break;
            case 2: strobe(15, false); break;   // FAST_DARK
            case 3: strobe(35, false); break;   // SLOW_DARK
            case 4: strobe(15, false); break;
            case 8:  // lights.rs:177-189
                t.kind = GLOWING_LIGHT;
                t.min_light = find_min_surrounding_light(map, i, sec.light_level);
                t.max_light = sec.light_level;
                t.going_up = false;
                list.push_back(t);
                break;
            case 12: strobe(35, true); break;
            case 13: strobe(15, true); break;
            case 17:  // lights.rs:225-239
                t.kind = FIRE_FLICKER;
                t.min_light = (int16_t)(find_min_surrounding_light(map, i, sec.light_level) + 16);
                t.max_light = sec.light_level;
                t.count = 4;
                list.push_back(t);
                break;
            default: break;
            }
        }
        for (int i = 0; i < (int)map.objects.size(); i++) {  // map_objects.rs:69-74
            T t{};
            t.kind = MAP_OBJECT;
            t.target = i;
            t.count = DRR_STATES[map.objects[i].state].tics;
            list.push_back(t);
        }
    }

    static void move_to_state(MapObject &mo, T &t, int state) {  // map_objects.rs:76-82
        mo.state = state;
        const DrrState &st = DRR_STATES[state];
        mo.sprite = st.sprite;
        mo.frame = st.frame;
        mo.full_bright = st.full_bright != 0;
        mo.is_null = state == 0;  // renderer/map_objects.rs:37
        t.count = st.tics;
    }

    void tick(Map &map, Pcg32 &rng) {  // game.rs:456-460: every thinker's mutate()
        for (T &t : list) {
            switch (t.kind) {
            case LIGHT_FLASH: {  // lights.rs:81-101
                Sector &sec = map.sectors[t.target];
                t.count = (int16_t)(t.count - 1);
                if (t.count > 0) break;
                if (sec.light_level == t.max_light) {
                    sec.light_level = t.min_light;
                    t.count = rng.gen_range(1, t.min_time + 1);
                } else {
                    sec.light_level = t.max_light;
                    t.count = rng.gen_range(1, t.max_time + 1);
                }
                break;
            }
            case STROBE_FLASH: {  // lights.rs:144-164
                Sector &sec = map.sectors[t.target];
                t.count = (int16_t)(t.count - 1);
                if (t.count > 0) break;
                if (sec.light_level == t.max_light) {
                    sec.light_level = t.min_light;
                    t.count = t.dark_time;
                } else {
                    sec.light_level = t.max_light;
                    t.count = t.bright_time;
                }
                break;
            }
            case GLOWING_LIGHT: {  // lights.rs:192-212 (GLOW_SPEED 8)
                Sector &sec = map.sectors[t.target];
                if (t.going_up) {
                    sec.light_level = (int16_t)(sec.light_level + 8);
                    if (sec.light_level >= t.max_light) {
                        sec.light_level = (int16_t)(sec.light_level - 8);
                        t.going_up = false;
                    }
                } else {
                    sec.light_level = (int16_t)(sec.light_level - 8);
                    if (sec.light_level <= t.min_light) {
                        sec.light_level = (int16_t)(sec.light_level + 8);
                        t.going_up = true;
                    }
                }
                break;
            }
            case FIRE_FLICKER: {  // lights.rs:242-259
                Sector &sec = map.sectors[t.target];
                t.count = (int16_t)(t.count - 1);
                if (t.count > 0) break;
                const int16_t amount = (int16_t)(rng.gen_range(0, 4) * 16);
                if ((int16_t)(sec.light_level - amount) < t.min_light)
                    sec.light_level = t.min_light;
                else
                    sec.light_level = (int16_t)(t.max_light - amount);
                t.count = 4;
                break;
            }
            case MAP_OBJECT: {  // map_objects.rs:84-95
                if (t.count == -1) break;
                t.count = (int16_t)(t.count - 1);
                if (t.count > 0) break;
                MapObject &mo = map.objects[t.target];
                move_to_state(mo, t, DRR_STATES[mo.state].next);
                break;
            }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// renderer
// ---------------------------------------------------------------------------------------
struct Player {  // game.rs:41-45
    Vertex position;
    float floor_height;
    float angle;
};
struct ClippedLine {  // renderer/clipped_line.rs
    Line line;
    float start_offset;
};
struct SdlLine {  // renderer/sdl_line.rs (sdl2::rect::Point == two i32)
    int32_t sx, sy, ex, ey;
};
struct Visplane {  // renderer/visplanes.rs:17-40
    std::shared_ptr<Flat> flat;
    int16_t height, light_level, left = -1, right = -1;
    std::vector<int16_t> top, bottom;
    Visplane(const std::shared_ptr<Flat> &f, int16_t h, int16_t l, int W) : flat(f), height(h), light_level(l), top(W, 0), bottom(W, 0) {}
};
struct BitmapColumn {  // bitmap_render.rs:19-25
    int32_t x, clipped_top_y, clipped_bottom_y, bottom_y, top_y;
};
enum RenderState { SolidSeg, TwoSidedSeg, DrawnSeg, MapObjectState };  // bitmap_render.rs:12-17
struct BitmapRender {                                                  // bitmap_render.rs:29-45
    RenderState state;
    std::shared_ptr<Bitmap> bitmap;
    int16_t light_level;
    ClippedLine clipped_line;
    int32_t start_x, end_x;
    float bottom_height, top_height;
    int16_t offset_x, offset_y;
    bool extends_to_bottom, extends_to_top, draw_ceiling;
    std::vector<BitmapColumn> columns;
};

// One recorded leaf call (used by the replay / kernel-level parity tests).
struct TraceCall {
    int kind;  // 0 = render_vertical_bitmap_line, 1 = draw_visplane
    int phase; // 0 = A (segs.rs:234), 1 = B (mod.rs:108), 2 = C/D (bitmap_render.rs:109)
    // kind 0
    int bitmap_id;
    int16_t light_level;
    ClippedLine cl;
    int32_t start_x, end_x;
    float bottom_height, top_height;
    int16_t offset_x, offset_y;
    int32_t x, clipped_bottom_y, clipped_top_y, bottom_y, top_y;
    // kind 1
    int flat_id;
    int is_sky;
    int16_t height, left, right;
    std::vector<int16_t> top, bottom;  // full W-sized arrays, untouched
};

enum { PHASE_WALLS = 1, PHASE_PLANES = 2, PHASE_MASKED = 4 };

struct Frame {
    int W, H;
    // constants.rs:3-17, derived from W,H exactly as the reference derives them from SCREEN_WIDTH/HEIGHT
    float ASPECT, GSW, GCFX, CFX, CFY;
    uint8_t *pixels = nullptr;  // pixels.rs:5-14: W*H*3, zeroed by the caller per frame
    int phases = 7;
    std::vector<TraceCall> *trace = nullptr;

    Frame(int w, int h) : W(w), H(h) {
        ASPECT = 200.0f / 240.0f;
        GSW = (float)(uint32_t)W / ASPECT;
        GCFX = GSW / 2.0f;
        CFX = (float)(uint32_t)W / 2.0f;
        CFY = (float)(uint32_t)H / 2.0f;
    }
    void set(uint64_t x, uint64_t y, const Color &c) {  // pixels.rs:22-30
        if (x >= (uint64_t)W || y > (uint64_t)H) return;
        if (y == (uint64_t)H) rs_panic("Pixels::set index out of bounds (y == SCREEN_HEIGHT)");
        uint8_t *p = pixels + 3 * (y * (uint64_t)W + x);
        p[0] = c.r;
        p[1] = c.g;
        p[2] = c.b;
    }
};

// bitmap_render.rs:190-208
static Color diminish_color(const Color &color, int16_t light_level, int16_t distance) {
    float factor = (float)light_level / 255.0f;
    const float dimishing_factor = 1.0f / (16.0f * 256.0f);
    factor -= (float)distance * dimishing_factor;
    if (factor < 0.0f) factor = 0.0f;
    return {f2u8((float)color.r * factor), f2u8((float)color.g * factor), f2u8((float)color.b * factor)};
}

// bitmap_render.rs:213-276
static void render_vertical_bitmap_line(Frame &fr, const Assets &as, const Bitmap &bitmap, int16_t light_level,
                                        const ClippedLine &clipped_line, int32_t start_x, int32_t end_x, float bottom_height,
                                        float top_height, int16_t offset_x, int16_t offset_y, int32_t x, int32_t clipped_bottom_y,
                                        int32_t clipped_top_y, int32_t bottom_y, int32_t top_y) {
    float len = line_length(clipped_line.line);
    float ux0 = 0.0f, ux1 = len;
    float uy0 = 0.0f, uy1 = top_height - bottom_height;
    float uz0 = clipped_line.line.start.x, uz1 = clipped_line.line.end.x;

    float ax = (float)sub32(x, start_x) / (float)sub32(end_x, start_x);
    if (bitmap.width == 0) rs_panic("attempt to divide by zero (bitmap.width)");
    int16_t tx = f2i16(((1.0f - ax) * (ux0 / uz0) + ax * (ux1 / uz1)) / ((1.0f - ax) * (1.0f / uz0) + ax * (1.0f / uz1)));
    tx = add16(tx, add16(f2i16(clipped_line.start_offset), offset_x));
    if (tx < 0) tx = add16(tx, mul16(bitmap.width, sub16(1, (int16_t)(tx / bitmap.width))));
    tx = (int16_t)(tx % bitmap.width);

    int16_t z = f2i16(((1.0f - ax) + ax) / ((1.0f - ax) * (1.0f / uz0) + ax * (1.0f / uz1)));

    for (int32_t y = clipped_top_y; y < clipped_bottom_y + 1; y++) {
        float ay = (float)sub32(y, top_y) / (float)sub32(bottom_y, top_y);
        if (bitmap.height == 0) rs_panic("attempt to divide by zero (bitmap.height)");
        int16_t ty = f2i16((float)bitmap.height + (1.0f - ay) * uy0 + ay * uy1);
        ty = add16(ty, offset_y);
        if (ty < 0) ty = add16(ty, mul16(bitmap.height, sub16(1, (int16_t)(ty / bitmap.height))));
        ty = (int16_t)(ty % bitmap.height);

        int16_t texel = bitmap.get(ty, tx);  // `as usize` of a negative index panics in Rust -> bounds error here
        if (texel >= 0) {
            Color color = as.palette[texel];
            Color dim = diminish_color(color, light_level, z);
            fr.set((uint64_t)(int64_t)x, (uint64_t)(int64_t)y, dim);
        }
    }
}

// visplanes.rs:42-80
static void draw_sky(Frame &fr, const Assets &as, const Player &player, const Bitmap &sky, const Visplane &vp) {
    const int16_t SKY_W = 256, SKY_H = 128;
    int16_t tx_offset = add16(f2i16((float)(int16_t)(-SKY_W) * player.angle / (PI_F / 2.0f)), SKY_W);
    if (tx_offset < 0) tx_offset = add16(tx_offset, mul16(SKY_W, sub16(1, (int16_t)(tx_offset / SKY_W))));

    for (int16_t x = vp.left; x < add16(vp.right, 1); x++) {
        if (x < 0 || x >= fr.W) rs_panic("visplane x out of bounds");
        int16_t top = std::max<int16_t>(vp.top[x], 0);
        int16_t bottom = std::min<int16_t>(vp.bottom[x], (int16_t)(fr.H - 1));
        for (int16_t y = top; y < add16(bottom, 1); y++) {
            int16_t tx = f2i16((float)x * (float)SKY_W / (float)(uint32_t)fr.W);
            tx = (int16_t)(add16(tx, tx_offset) % SKY_W);
            int16_t ty = f2i16((float)y * (float)SKY_H * 2.0f / (float)(uint32_t)fr.H);
            if (ty < 0) ty = add16(ty, SKY_H);
            ty = (int16_t)(ty % SKY_H);
            int16_t texel = sky.get(ty, tx);
            if (texel >= 0) fr.set((uint64_t)(int64_t)x, (uint64_t)(int64_t)y, as.palette[texel]);
        }
    }
}

// visplanes.rs:82-130
static void draw_visplane(Frame &fr, const Assets &as, const Player &player, const Bitmap &sky, const Visplane &vp) {
    if (vp.flat->name.find("SKY") != std::string::npos) {
        draw_sky(fr, as, player, sky, vp);
        return;
    }
    const int16_t FLAT_SIZE = 64;
    for (int16_t x = vp.left; x < add16(vp.right, 1); x++) {
        if (x < 0 || x >= fr.W) rs_panic("visplane x out of bounds");
        int16_t top = std::max<int16_t>(vp.top[x], 0);
        int16_t bottom = std::min<int16_t>(vp.bottom[x], (int16_t)(fr.H - 1));
        if (sub16(bottom, top) <= 1) continue;
        for (int16_t y = top; y < add16(bottom, 1); y++) {
            float vx = (fr.CFX - (float)x) / fr.ASPECT;
            float vy = fr.CFY - (float)y;
            float wz = (float)vp.height - player.floor_height - 41.0f;
            float wx = fr.GCFX * wz / vy;
            float wy = wz * vx / vy;
            Vertex rotated = rotate({wx, wy}, player.angle);
            int16_t tx = add16(f2i16(rotated.x), f2i16(player.position.x));
            int16_t ty = add16(f2i16(rotated.y), f2i16(player.position.y));
            tx &= (int16_t)(FLAT_SIZE - 1);
            ty &= (int16_t)(FLAT_SIZE - 1);
            Color color = as.palette[vp.flat->pixels[ty][tx]];
            Color dim = diminish_color(color, vp.light_level, f2i16(wx));
            fr.set((uint64_t)(int64_t)x, (uint64_t)(int64_t)y, dim);
        }
    }
}

// misc.rs:13-115
static bool clip_to_viewport(const Line &line, ClippedLine *out) {
    Line left = {{0.0f, 0.0f}, {1.0f, 1.0f}};
    Line right = {{0.0f, 0.0f}, {1.0f, -1.0f}};
    bool start_outside_left = is_left_of_line(line.start, left);
    bool end_outside_left = is_left_of_line(line.end, left);
    bool start_outside_right = !is_left_of_line(line.start, right);
    bool end_outside_right = !is_left_of_line(line.end, right);
    bool start_in_viewport = line.start.x > 0.0f && !start_outside_left && !start_outside_right;
    bool end_in_viewport = line.end.x > 0.0f && !end_outside_left && !end_outside_right;
    if (start_in_viewport && end_in_viewport) {
        *out = {line, 0.0f};
        return true;
    }
    Vertex li, ri;
    bool l_ok = intersection(line, left, &li), r_ok = intersection(line, right, &ri);
    bool left_intersected = l_ok ? li.x >= 0.0f : false;
    bool right_intersected = r_ok ? ri.x >= 0.0f : false;
    if (!start_in_viewport && !end_in_viewport && !left_intersected && !right_intersected) return false;
    if (!start_in_viewport && !end_in_viewport && (left_intersected != right_intersected)) return false;
    if ((right_intersected && start_outside_right && end_outside_right) || (left_intersected && start_outside_left && end_outside_left))
        return false;
    float start_offset = 0.0f;
    Vertex start = line.start, end = line.end;
    if (left_intersected) {
        if (start_outside_left) {
            Vertex new_start = li;
            start_offset = distance_to(new_start, start);
            start = new_start;
        }
        if (end_outside_left) end = li;
    }
    if (right_intersected) {
        if (start_outside_right) start = ri;
        if (end_outside_right) end = ri;
    }
    *out = {{start, end}, start_offset};
    return true;
}

// misc.rs:130-161
static SdlLine make_sidedef_non_vertical_line(const Frame &fr, const Line &line, float height) {
    auto persp = [&](const Vertex &v, float y) -> Vertex {
        float x = v.y, z = v.x;
        return {fr.GCFX * x / z, fr.GCFX * y / z};
    };
    Vertex ts = persp(line.start, height), te = persp(line.end, height);
    ts.x *= fr.ASPECT;
    te.x *= fr.ASPECT;
    SdlLine l;
    l.sx = f2i32(fr.CFX - ts.x);
    l.sy = f2i32(fr.CFY - ts.y);
    l.ex = f2i32(fr.CFX - te.x);
    l.ey = f2i32(fr.CFY - te.y);
    l.sx = std::min(l.sx, fr.W - 1);
    l.ex = std::min(l.ex, fr.W - 1);
    return l;
}

// sidedef_visplanes.rs
struct SidedefVisPlanes {
    int W;
    int16_t light_level;
    std::shared_ptr<Flat> floor_flat, ceiling_flat;
    int16_t floor_height, ceiling_height;
    Visplane bottom_visplane, top_visplane;
    bool bottom_used = false, top_used = false;
    SidedefVisPlanes(int W_, int16_t l, const std::shared_ptr<Flat> &ff, const std::shared_ptr<Flat> &cf, int16_t fh, int16_t ch)
        : W(W_), light_level(l), floor_flat(ff), ceiling_flat(cf), floor_height(fh), ceiling_height(ch),
          bottom_visplane(ff, fh, l, W_), top_visplane(cf, ch, l, W_) {}
    void flush(std::vector<Visplane> &out) {  // :41-58
        if (bottom_used) {
            out.push_back(bottom_visplane);
            bottom_visplane = Visplane(floor_flat, floor_height, light_level, W);
            bottom_used = false;
        }
        if (top_used) {
            out.push_back(top_visplane);
            top_visplane = Visplane(ceiling_flat, ceiling_height, light_level, W);
            top_used = false;
        }
    }
    void add_bottom_point(int16_t x, int16_t top_y, int16_t bottom_y) {  // :60-71
        if (!bottom_used) bottom_visplane.left = x;
        bottom_visplane.right = x;
        bottom_used = true;
        bottom_visplane.top[x] = top_y;
        bottom_visplane.bottom[x] = bottom_y;
    }
    void add_top_point(int16_t x, int16_t top_y, int16_t bottom_y) {  // :73-84
        if (!top_used) top_visplane.left = x;
        top_visplane.right = x;
        top_used = true;
        top_visplane.top[x] = top_y;
        top_visplane.bottom[x] = bottom_y;
    }
};

struct Game;

struct Renderer {  // renderer/mod.rs + segs.rs
    Frame &fr;
    Assets &as;
    const Map &map;
    const Player &player;
    std::shared_ptr<Bitmap> sky;
    float timestamp;

    std::vector<BitmapRender> segs;
    std::vector<Visplane> visplanes;
    std::vector<uint8_t> hor_ocl;
    std::vector<int16_t> floor_ver_ocl, ceiling_ver_ocl;

    Renderer(Frame &f, Assets &a, const Map &m, const Player &p, std::shared_ptr<Bitmap> s, float ts)
        : fr(f), as(a), map(m), player(p), sky(s), timestamp(ts), hor_ocl(f.W, 0), floor_ver_ocl(f.W, (int16_t)f.H),
          ceiling_ver_ocl(f.W, -1) {}  // segs.rs:80-101

    void trace_column(int phase, const Bitmap &bm, int16_t light, const ClippedLine &cl, int32_t sx, int32_t ex, float bh, float th,
                      int16_t ox, int16_t oy, const BitmapColumn &c) {
        if (!fr.trace) return;
        TraceCall t{};
        t.kind = 0;
        t.phase = phase;
        t.bitmap_id = bm.id;
        t.light_level = light;
        t.cl = cl;
        t.start_x = sx;
        t.end_x = ex;
        t.bottom_height = bh;
        t.top_height = th;
        t.offset_x = ox;
        t.offset_y = oy;
        t.x = c.x;
        t.clipped_bottom_y = c.clipped_bottom_y;
        t.clipped_top_y = c.clipped_top_y;
        t.bottom_y = c.bottom_y;
        t.top_y = c.top_y;
        fr.trace->push_back(t);
    }

    void occlude_vertical_line(int16_t x) {  // segs.rs:113-117
        hor_ocl[x] = 1;
        floor_ver_ocl[x] = (int16_t)(fr.H) / 2;
        ceiling_ver_ocl[x] = (int16_t)(fr.H) / 2;
    }

    struct SideDefDetails {
        const ClippedLine *clipped_line;
        const Sidedef *sidedef;
        int16_t offset_x, floor_height, ceiling_height;
        std::shared_ptr<Flat> floor_flat, ceiling_flat;
        int16_t light_level;
    };
    struct Flags {
        bool only_occlusions, is_lower_wall, is_upper_wall, draw_ceiling, is_two_sided_middle_wall;
    };

    // segs.rs:121-350
    void process_sidedef(const SideDefDetails &sds, float bottom_height, float top_height, int32_t offset_y,
                         const std::string &texture_name, Flags flags) {
        SdlLine bottom = make_sidedef_non_vertical_line(fr, sds.clipped_line->line, bottom_height);
        SdlLine top = make_sidedef_non_vertical_line(fr, sds.clipped_line->line, top_height);

        std::shared_ptr<Bitmap> texture;
        if (texture_name != "-") texture = as.texture_get(texture_name);

        if (bottom.sx != top.sx || bottom.ex != top.ex) rs_panic("Wall start not vertical");
        if ((int16_t)bottom.sx == (int16_t)bottom.ex || (int16_t)top.sx == (int16_t)top.ex) return;
        auto check = [&](const SdlLine &l) {  // :103-111
            if (l.sx < 0 || l.sx >= fr.W) rs_panic("Invalid line start x: " + std::to_string(l.sx));
            if (l.ex < 0 || l.ex >= fr.W) rs_panic("Invalid line end x: " + std::to_string(l.ex));
        };
        check(bottom);
        check(top);

        float bottom_delta = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
        float top_delta = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);

        SidedefVisPlanes sv(fr.W, sds.light_level, sds.floor_flat, sds.ceiling_flat, sds.floor_height, sds.ceiling_height);

        bool is_full_height_wall = !flags.is_lower_wall && !flags.is_upper_wall && !flags.only_occlusions;
        RenderState st = flags.is_two_sided_middle_wall ? TwoSidedSeg : SolidSeg;

        int16_t off_x = add16(f2i16(sds.sidedef->x_offset), sds.offset_x);
        int16_t off_y = add16(f2i16(sds.sidedef->y_offset), w16(offset_y));

        BitmapRender br;
        br.state = st;
        br.bitmap = texture;
        br.light_level = sds.light_level;
        br.clipped_line = *sds.clipped_line;
        br.start_x = bottom.sx;
        br.end_x = bottom.ex;
        br.bottom_height = bottom_height;
        br.top_height = top_height;
        br.offset_x = off_x;
        br.offset_y = off_y;
        br.extends_to_bottom = flags.is_lower_wall || (!flags.is_two_sided_middle_wall && is_full_height_wall);
        br.extends_to_top = flags.is_upper_wall || (!flags.is_two_sided_middle_wall && is_full_height_wall);
        br.draw_ceiling = flags.draw_ceiling;

        const int16_t H16 = (int16_t)fr.H;
        for (int16_t x = (int16_t)bottom.sx; x < add16((int16_t)bottom.ex, 1); x++) {
            if (!hor_ocl[x]) {
                int16_t bottom_y = f2i16((float)bottom.sy + ((float)x - (float)bottom.sx) * bottom_delta);
                int16_t top_y = f2i16((float)top.sy + ((float)x - (float)top.sx) * top_delta);

                int16_t fvo = floor_ver_ocl[x], cvo = ceiling_ver_ocl[x];
                int16_t clipped_bottom_y = std::min(fvo, bottom_y);
                int16_t clipped_top_y = std::max(cvo, top_y);
                clipped_bottom_y = std::min<int16_t>(sub16(H16, 1), clipped_bottom_y);
                clipped_top_y = std::max<int16_t>(0, clipped_top_y);

                bool in_ver_clipped_area = clipped_bottom_y >= clipped_top_y;

                if (in_ver_clipped_area) {
                    BitmapColumn col{x, clipped_top_y, clipped_bottom_y, bottom_y, top_y};
                    if (!flags.is_two_sided_middle_wall && !flags.only_occlusions) {
                        if (texture) {
                            trace_column(0, *texture, sds.light_level, *sds.clipped_line, bottom.sx, bottom.ex, bottom_height, top_height,
                                         off_x, off_y, col);
                            if (fr.phases & PHASE_WALLS)
                                render_vertical_bitmap_line(fr, as, *texture, sds.light_level, *sds.clipped_line, bottom.sx, bottom.ex,
                                                            bottom_height, top_height, off_x, off_y, x, clipped_bottom_y, clipped_top_y,
                                                            bottom_y, top_y);
                        }
                    }
                    br.columns.push_back(col);  // add_column, bitmap_render.rs:84-99
                }

                if (!flags.is_two_sided_middle_wall && in_ver_clipped_area && (is_full_height_wall || flags.only_occlusions)) {
                    bool visplane_added = false;
                    if (clipped_bottom_y < fvo && clipped_bottom_y != sub16(H16, 1)) {
                        sv.add_bottom_point(x, clipped_bottom_y, fvo);
                        visplane_added = true;
                    }
                    if (!flags.is_two_sided_middle_wall && flags.draw_ceiling && clipped_top_y > cvo && clipped_top_y != -1) {
                        if (flags.draw_ceiling) sv.add_top_point(x, cvo, clipped_top_y);
                        visplane_added = true;
                    }
                    if (!visplane_added) sv.flush(visplanes);
                } else if (!flags.is_two_sided_middle_wall && !in_ver_clipped_area && (is_full_height_wall || flags.only_occlusions) &&
                           fvo > cvo) {
                    if (bottom_y <= cvo) {
                        sv.add_bottom_point(x, cvo, fvo);
                        occlude_vertical_line(x);
                    }
                    if (flags.draw_ceiling && top_y >= fvo) {
                        if (flags.draw_ceiling) sv.add_top_point(x, cvo, fvo);
                        occlude_vertical_line(x);
                    }
                }

                if (!flags.is_two_sided_middle_wall && in_ver_clipped_area && flags.only_occlusions) {
                    floor_ver_ocl[x] = clipped_bottom_y;
                    if (flags.draw_ceiling) ceiling_ver_ocl[x] = clipped_top_y;
                }
                if (!flags.is_two_sided_middle_wall && in_ver_clipped_area && flags.is_lower_wall) floor_ver_ocl[x] = clipped_top_y;
                if (!flags.is_two_sided_middle_wall && in_ver_clipped_area && flags.is_upper_wall) ceiling_ver_ocl[x] = clipped_bottom_y;
            } else {
                sv.flush(visplanes);
            }
            if (!flags.is_two_sided_middle_wall && is_full_height_wall) occlude_vertical_line(x);
        }
        sv.flush(visplanes);
        segs.push_back(std::move(br));
    }

    // segs.rs:353-590
    void process_seg(const Seg &seg) {
        const Linedef &linedef = map.linedefs[seg.linedef];
        int front_i = seg.direction ? linedef.back_sidedef : linedef.front_sidedef;
        int back_i = seg.direction ? linedef.front_sidedef : linedef.back_sidedef;
        if (front_i == -1) return;
        const Sidedef &front_sidedef = map.sidedefs[front_i];
        const Sector &front_sector = map.sectors[front_sidedef.sector];

        float floor_height = (float)front_sector.floor_height;
        float ceiling_height = (float)front_sector.ceiling_height;

        bool has_pb = false, has_pt = false;
        float portal_bottom_height = 0.0f, portal_top_height = 0.0f;
        if (back_i != -1) {
            const Sector &back_sector = map.sectors[map.sidedefs[back_i].sector];
            if (back_sector.floor_height > front_sector.floor_height) {
                has_pb = true;
                portal_bottom_height = (float)back_sector.floor_height;
            }
            if (back_sector.ceiling_height < front_sector.ceiling_height) {
                has_pt = true;
                portal_top_height = (float)back_sector.ceiling_height;
            }
        }
        bool is_two_sided = (linedef.flags & 4) != 0;
        bool top_is_unpegged = (linedef.flags & 8) != 0;
        bool bottom_is_unpegged = (linedef.flags & 16) != 0;

        Vertex moved_start = vsub(map.vertexes[seg.start_vertex], player.position);
        Vertex moved_end = vsub(map.vertexes[seg.end_vertex], player.position);
        Vertex start = rotate(moved_start, -player.angle);
        Vertex end = rotate(moved_end, -player.angle);
        Line line = {start, end};

        ClippedLine clipped_line;
        if (!clip_to_viewport(line, &clipped_line)) return;
        if (clipped_line.line.start.x < -0.01f) rs_panic("Clipped line x < -0.01");

        float player_height = player.floor_height + 41.0f;
        SdlLine floor = make_sidedef_non_vertical_line(fr, clipped_line.line, floor_height - player_height);
        if (floor.sx > floor.ex) return;

        auto floor_flat = as.flat_get_animated(front_sector.floor_texture, timestamp);
        auto ceiling_flat = as.flat_get_animated(front_sector.ceiling_texture, timestamp);

        bool draw_ceiling = true;
        if (back_i != -1) {
            const Sector &back_sector = map.sectors[map.sidedefs[back_i].sector];
            if (front_sector.ceiling_texture.find("SKY") != std::string::npos &&
                back_sector.ceiling_texture.find("SKY") != std::string::npos) {
                float back_ceiling = (float)back_sector.ceiling_height;
                has_pt = false;
                ceiling_height = fmin_rs(back_ceiling, ceiling_height);
                draw_ceiling = false;
            }
        }

        SideDefDetails sds{&clipped_line,           &front_sidedef, seg.offset,  front_sector.floor_height,
                           front_sector.ceiling_height, floor_flat,     ceiling_flat, front_sector.light_level};

        if (!is_two_sided) {
            int32_t offset_y = bottom_is_unpegged ? f2i32(floor_height - ceiling_height) : 0;
            process_sidedef(sds, floor_height - player_height, ceiling_height - player_height, offset_y, front_sidedef.middle_texture,
                            {false, false, false, draw_ceiling, false});
        } else {
            process_sidedef(sds, floor_height - player_height, ceiling_height - player_height, 0, front_sidedef.middle_texture,
                            {true, false, false, draw_ceiling, false});
            float mid_floor = floor_height, mid_ceiling = ceiling_height;
            if (has_pb) mid_floor = portal_bottom_height;
            if (has_pt) mid_ceiling = portal_top_height;
            process_sidedef(sds, mid_floor - player_height, mid_ceiling - player_height, 0, front_sidedef.middle_texture,
                            {false, false, false, draw_ceiling, true});
            if (has_pb) {
                int32_t offset_y = bottom_is_unpegged ? f2i32(ceiling_height - portal_bottom_height) : 0;
                process_sidedef(sds, floor_height - player_height, portal_bottom_height - player_height, offset_y,
                                front_sidedef.lower_texture, {false, true, false, draw_ceiling, false});
            }
            if (has_pt) {
                int32_t offset_y = top_is_unpegged ? 0 : f2i32(portal_top_height - ceiling_height);
                process_sidedef(sds, portal_top_height - player_height, ceiling_height - player_height, offset_y,
                                front_sidedef.upper_texture, {false, false, true, draw_ceiling, false});
            }
        }
    }

    // mod.rs:61-104
    void process_subsector(const SubSector &ss) {
        for (int s : ss.segs) process_seg(map.segs[s]);
    }
    void render_node(const Node &node) {
        Vertex v1 = {node.x, node.y};
        Vertex v2 = vadd(v1, {node.dx, node.dy});
        bool is_left = is_left_of_line(player.position, {v1, v2});
        bool f_ss = is_left ? node.left_is_subsector : node.right_is_subsector;
        int f_ch = is_left ? node.left_child : node.right_child;
        bool b_ss = is_left ? node.right_is_subsector : node.left_is_subsector;
        int b_ch = is_left ? node.right_child : node.left_child;
        if (f_ss)
            process_subsector(map.subsectors[f_ch]);
        else
            render_node(map.nodes[f_ch]);
        if (b_ss)
            process_subsector(map.subsectors[b_ch]);
        else
            render_node(map.nodes[b_ch]);
    }

    // bitmap_render.rs:101-135
    void render_bitmap_render(BitmapRender &br) {
        if (br.state == SolidSeg || br.state == DrawnSeg) return;
        if (br.bitmap) {
            for (auto &c : br.columns) {
                trace_column(2, *br.bitmap, br.light_level, br.clipped_line, br.start_x, br.end_x, br.bottom_height, br.top_height,
                             br.offset_x, br.offset_y, c);
                if (fr.phases & PHASE_MASKED)
                    render_vertical_bitmap_line(fr, as, *br.bitmap, br.light_level, br.clipped_line, br.start_x, br.end_x,
                                                br.bottom_height, br.top_height, br.offset_x, br.offset_y, c.x, c.clipped_bottom_y,
                                                c.clipped_top_y, c.bottom_y, c.top_y);
            }
        }
        br.state = DrawnSeg;
    }
    // bitmap_render.rs:137-165
    static bool is_behind_vertex(const BitmapRender &br, const Vertex &v) {
        float min_x = fmin_rs(br.clipped_line.line.start.x, br.clipped_line.line.end.x);
        float max_x = fmax_rs(br.clipped_line.line.start.x, br.clipped_line.line.end.x);
        if (min_x > v.x) return true;
        if (max_x > v.x && !is_left_of_line(v, br.clipped_line.line)) return true;
        return false;
    }

    // renderer/bsp.rs:9-44 ; returns sector index or -1
    static int get_sector_from_vertex(const Map &map, const Vertex &v) {
        const Node *node = &map.nodes.back();
        for (;;) {
            Vertex v1 = {node->x, node->y};
            Vertex v2 = vadd(v1, {node->dx, node->dy});
            bool is_left = is_left_of_line(v, {v1, v2});
            bool ss = is_left ? node->left_is_subsector : node->right_is_subsector;
            int ch = is_left ? node->left_child : node->right_child;
            if (!ss) {
                node = &map.nodes[ch];
                continue;
            }
            for (int s : map.subsectors[ch].segs) {
                const Seg &seg = map.segs[s];
                const Linedef &ld = map.linedefs[seg.linedef];
                int sd = seg.direction ? ld.back_sidedef : ld.front_sidedef;
                if (sd != -1) return map.sidedefs[sd].sector;
            }
            return -1;
        }
    }

    // renderer/map_objects.rs:19-241
    void draw_map_objects() {
        std::vector<BitmapRender> mo_renders;
        const int16_t H16 = (int16_t)fr.H;
        for (const MapObject &mo : map.objects) {
            if (mo.is_null) continue;
            float angle = player.angle - mo.angle - PI_F;
            angle += PI_F / 16.0f;
            angle = fmodf(angle, 2.0f * PI_F);
            if (angle < 0.0f) angle += 2.0f * PI_F;
            angle = fmodf(angle, 2.0f * PI_F);
            uint8_t rotation = f2u8(angle * 8.0f / (2.0f * PI_F));
            auto picture = as.sprite_get_picture(mo.sprite, mo.frame, rotation);

            Vertex moved = vsub(mo.position, player.position);
            Vertex vpv = rotate(moved, -player.angle);
            int16_t width = picture->bitmap->width;
            Vertex start = vsub(vpv, {0.0f, (float)(int16_t)(-width) / 2.0f});
            Vertex end = vsub(vpv, {0.0f, (float)width / 2.0f});
            Line line = {start, end};
            ClippedLine clipped_line;
            if (!clip_to_viewport(line, &clipped_line)) continue;
            if (clipped_line.line.start.x < -0.01f) rs_panic("Clipped line x < -0.01 (map object)");

            int sector = get_sector_from_vertex(map, mo.position);
            if (sector < 0) continue;  // "Thing is outside map"
            int16_t light_level = mo.full_bright ? (int16_t)255 : map.sectors[sector].light_level;

            float player_height = player.floor_height + 41.0f;
            int16_t z = map.sectors[sector].floor_height;
            float bottom_height = (float)z - player_height;
            float top_height = (float)z + (float)picture->bitmap->height - 1.0f - player_height;
            bottom_height += (float)picture->top_offset - (float)picture->bitmap->height;
            top_height += (float)picture->top_offset - (float)picture->bitmap->height;

            SdlLine bottom = make_sidedef_non_vertical_line(fr, clipped_line.line, bottom_height);
            SdlLine top = make_sidedef_non_vertical_line(fr, clipped_line.line, top_height);

            std::vector<int16_t> top_seg_clip(fr.W, -1), bottom_seg_clip(fr.W, H16);
            for (auto &seg : segs) {
                if (is_behind_vertex(seg, vpv)) continue;
                for (auto &column : seg.columns) {
                    if (column.x < 0 || column.x >= fr.W) rs_panic("seg column x out of bounds");
                    size_t x = (size_t)column.x;
                    if (seg.state == SolidSeg) {
                        if (seg.extends_to_bottom) bottom_seg_clip[x] = std::min(bottom_seg_clip[x], (int16_t)column.clipped_top_y);
                        if (seg.extends_to_top) top_seg_clip[x] = std::max(top_seg_clip[x], (int16_t)column.clipped_bottom_y);
                    } else if (seg.state == TwoSidedSeg) {
                        if (seg.draw_ceiling) top_seg_clip[x] = std::max(top_seg_clip[x], (int16_t)column.top_y);
                        bottom_seg_clip[x] = std::min(bottom_seg_clip[x], (int16_t)column.bottom_y);
                    }
                }
            }

            BitmapRender br;
            br.state = MapObjectState;
            br.bitmap = picture->bitmap;
            br.light_level = light_level;
            br.clipped_line = clipped_line;
            br.start_x = bottom.sx;
            br.end_x = bottom.ex;
            br.bottom_height = bottom_height;
            br.top_height = top_height;
            br.offset_x = 0;
            br.offset_y = 0;
            br.extends_to_bottom = br.extends_to_top = br.draw_ceiling = false;

            float bottom_delta = ((float)bottom.sy - (float)bottom.ey) / ((float)bottom.sx - (float)bottom.ex);
            float top_delta = ((float)top.sy - (float)top.ey) / ((float)top.sx - (float)top.ex);
            for (int16_t x = (int16_t)bottom.sx; x < (int16_t)bottom.ex; x++) {
                int16_t bottom_y = f2i16((float)bottom.sy + ((float)x - (float)bottom.sx) * bottom_delta);
                int16_t top_y = f2i16((float)top.sy + ((float)x - (float)top.sx) * top_delta);
                if (x < 0 || x >= fr.W) rs_panic("map object column x out of bounds");
                int16_t clipped_top_y = std::max(top_y, top_seg_clip[x]);
                int16_t clipped_bottom_y = std::min(bottom_y, bottom_seg_clip[x]);
                clipped_top_y = std::max<int16_t>(0, clipped_top_y);
                clipped_bottom_y = std::min<int16_t>(sub16(H16, 1), clipped_bottom_y);
                br.columns.push_back({x, clipped_top_y, clipped_bottom_y, bottom_y, top_y});
            }
            mo_renders.push_back(std::move(br));
        }

        // sort() is a stable sort on `start.x as i16` (bitmap_render.rs:168-174), then reverse()
        std::stable_sort(mo_renders.begin(), mo_renders.end(), [](const BitmapRender &a, const BitmapRender &b) {
            return f2i16(a.clipped_line.line.start.x) < f2i16(b.clipped_line.line.start.x);
        });
        std::reverse(mo_renders.begin(), mo_renders.end());

        for (auto &mor : mo_renders) {
            Vertex v = {(mor.clipped_line.line.start.x + mor.clipped_line.line.end.x) / 2.0f,
                        (mor.clipped_line.line.start.y + mor.clipped_line.line.end.y) / 2.0f};
            for (auto &seg : segs)
                if (is_behind_vertex(seg, v)) render_bitmap_render(seg);
            render_bitmap_render(mor);
        }
    }

    // mod.rs:118-136
    void render() {
        render_node(map.nodes.back());
        for (auto &vp : visplanes) {  // draw_visplanes, mod.rs:106-116
            if (fr.trace) {
                TraceCall t{};
                t.kind = 1;
                t.phase = 1;
                t.flat_id = vp.flat->id;
                t.is_sky = vp.flat->name.find("SKY") != std::string::npos;
                t.height = vp.height;
                t.light_level = vp.light_level;
                t.left = vp.left;
                t.right = vp.right;
                t.top = vp.top;
                t.bottom = vp.bottom;
                fr.trace->push_back(std::move(t));
            }
            if (fr.phases & PHASE_PLANES) draw_visplane(fr, as, player, *sky, vp);
        }
        std::reverse(segs.begin(), segs.end());
        draw_map_objects();
        for (auto &s : segs) render_bitmap_render(s);  // draw_remaining_segs, segs.rs:593-597
    }
};

// game.rs:118-227 (headless part only)
struct Game {
    std::shared_ptr<WadFile> wad;
    Assets as;
    Map map;
    std::string map_name;
    int W, H;
    std::shared_ptr<Bitmap> sky;
    std::vector<TraceCall> trace;
    std::string last_error;

    static std::string sky_name(const std::string &map_name) {  // game.rs:199-227
        // regex e(\d+)m(\d+) searched anywhere in the (unmodified, case-sensitive) map name
        for (size_t i = 0; i + 1 < map_name.size(); i++) {
            if (map_name[i] != 'e') continue;
            size_t j = i + 1, d0 = j;
            while (j < map_name.size() && isdigit((unsigned char)map_name[j])) j++;
            if (j == d0 || j >= map_name.size() || map_name[j] != 'm') continue;
            size_t k = j + 1;
            if (k >= map_name.size() || !isdigit((unsigned char)map_name[k])) continue;
            long episode = strtol(map_name.substr(d0, j - d0).c_str(), nullptr, 10);
            if (episode == 2) return "SKY2";
            if (episode == 3) return "SKY3";
            return "SKY1";
        }
        for (size_t i = 0; i + 1 < map_name.size(); i++) {
            if (isdigit((unsigned char)map_name[i]) && isdigit((unsigned char)map_name[i + 1])) {
                int m = (map_name[i] - '0') * 10 + (map_name[i + 1] - '0');
                if (m < 12) return "SKY1";
                if (m < 21) return "SKY2";
                return "SKY3";
            }
        }
        return "SKY1";
    }

    std::vector<int16_t> light0;      // the WAD's sector light levels (tic 0)
    std::vector<MapObject> objects0;  // the spawn states (tic 0)

    // game.rs:462-482: `tic` game ticks from the start of the game (thinkers created by init_thinkers, game.rs:193)
    void set_tic(uint32_t tic, uint64_t seed) {
        for (size_t i = 0; i < map.sectors.size(); i++) map.sectors[i].light_level = light0[i];
        map.objects = objects0;
        Pcg32 rng(seed);
        Thinkers th;
        th.init(map, rng);
        for (uint32_t k = 0; k < tic; k++) th.tick(map, rng);
    }

    void init(std::shared_ptr<WadFile> w, const std::string &name, int W_, int H_) {
        wad = w;
        map_name = name;
        W = W_;
        H = H_;
        if (W <= 0 || H <= 0 || W > 32767 || H > 32767) rs_panic("bad screen size");
        map.load(*wad, name);
        for (auto &sec : map.sectors) light0.push_back(sec.light_level);
        objects0 = map.objects;
        as.wad = wad.get();
        as.load_palette();
        as.init_flats();
        as.init_textures();
        sky = as.texture_get(sky_name(name));
        as.init_sprites();
    }

    Player make_player(float x, float y, float angle) {  // game.rs:144-150, 376-389
        Player p{{x, y}, 0.0f, angle};
        int s = Renderer::get_sector_from_vertex(map, p.position);
        if (s >= 0) p.floor_height = (float)map.sectors[s].floor_height;
        return p;
    }

    void render(const Player &p, float timestamp, uint8_t *out, int phases, bool want_trace) {
        Frame fr(W, H);
        fr.pixels = out;
        fr.phases = phases;
        trace.clear();
        fr.trace = want_trace ? &trace : nullptr;
        memset(out, 0, (size_t)W * H * 3);  // Pixels::new, pixels.rs:10-14
        Renderer r(fr, as, map, p, sky, timestamp);
        r.render();
    }
};

}  // namespace orc

// ---------------------------------------------------------------------------------------
// C entry points for ctypes (tests/, smoke(), bench cpu_baseline only)
// ---------------------------------------------------------------------------------------
using namespace orc;
static thread_local std::string g_err;

template <class F>
static int guarded(F &&body) {
    try {
        body();
        return 0;
    } catch (const Error &e) {
        g_err = e.msg;
        return -1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -2;
    }
}

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

void *orc_game_new(const char *wad_path, const char *map_name, int W, int H) {
    try {
        FILE *f = fopen(wad_path, "rb");
        if (!f) {
            g_err = std::string("cannot open ") + wad_path;
            return nullptr;
        }
        std::vector<uint8_t> bytes;
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
        fclose(f);
        auto w = std::make_shared<WadFile>();
        w->load(std::move(bytes));
        auto *g = new Game();
        try {
            g->init(w, map_name, W, H);
        } catch (...) {
            delete g;
            throw;
        }
        return g;
    } catch (const Error &e) {
        g_err = e.msg;
    } catch (const std::exception &e) {
        g_err = e.what();
    }
    return nullptr;
}
void orc_game_free(void *g) { delete (Game *)g; }

// Player1Start (game.rs:151-157): out = {x, y, angle}
int orc_player_start(void *gp, float *out3) {
    Game *g = (Game *)gp;
    for (auto &t : g->map.things)
        if (t.thing_type == 1) {
            out3[0] = t.x;
            out3[1] = t.y;
            out3[2] = t.angle;
            return 0;
        }
    g_err = "Could not find thing of type 1";
    return -1;
}

// floor height the reference would assign to a player at (x,y) (game.rs:376-389); 0.0 if outside
float orc_floor_height_at(void *gp, float x, float y) { return ((Game *)gp)->make_player(x, y, 0.0f).floor_height; }
int orc_sector_at(void *gp, float x, float y) { return Renderer::get_sector_from_vertex(((Game *)gp)->map, {x, y}); }

// Render one frame exactly as Renderer::new(...).render() would (game.rs:505-519) into out (W*H*3, RGB24).
// phases: bit0 walls (segs.rs:234), bit1 visplanes (mod.rs:108), bit2 masked+sprites (bitmap_render.rs:109).
// Returns 0, or <0 if the reference would have panicked (message in orc_last_error()).
int orc_render(void *gp, float x, float y, float angle, float timestamp, uint8_t *out, int phases, int want_trace) {
    Game *g = (Game *)gp;
    return guarded([&] {
        Player p = g->make_player(x, y, angle);
        g->render(p, timestamp, out, phases, want_trace != 0);
    });
}

// The world `tic` game ticks (1/35 s, game.rs:32) after the start: sector light effects and map-object animation stepped
// by the reference's thinkers, random draws from a PCG32 stream seeded with `seed` (see Thinkers).  tic 0 = the WAD.
int orc_set_tic(void *gp, uint32_t tic, uint64_t seed) {
    Game *g = (Game *)gp;
    return guarded([&] { g->set_tic(tic, seed); });
}
// sector light levels / map object (sprite, frame, full_bright, is_null) quadruples of the current tic (for the tests)
int orc_sector_lights(void *gp, int16_t *out) {
    Game *g = (Game *)gp;
    for (size_t i = 0; i < g->map.sectors.size(); i++) out[i] = g->map.sectors[i].light_level;
    return (int)g->map.sectors.size();
}
int orc_object_states(void *gp, int32_t *out4) {
    Game *g = (Game *)gp;
    for (size_t i = 0; i < g->map.objects.size(); i++) {
        const MapObject &mo = g->map.objects[i];
        out4[4 * i] = mo.sprite;
        out4[4 * i + 1] = mo.frame;
        out4[4 * i + 2] = mo.full_bright;
        out4[4 * i + 3] = mo.is_null;
    }
    return (int)g->map.objects.size();
}

// ---- asset export (so a test can upload the very same assets through the C ABI) ----
int orc_palette(void *gp, uint8_t *out768) {
    Game *g = (Game *)gp;
    for (int i = 0; i < 256; i++) {
        out768[i * 3] = g->as.palette[i].r;
        out768[i * 3 + 1] = g->as.palette[i].g;
        out768[i * 3 + 2] = g->as.palette[i].b;
    }
    return 0;
}
int orc_bitmap_count(void *gp) { return (int)((Game *)gp)->as.bitmaps.size(); }
int orc_bitmap_size(void *gp, int id, int *w, int *h) {
    Game *g = (Game *)gp;
    if (id < 0 || id >= (int)g->as.bitmaps.size()) return -1;
    *w = g->as.bitmaps[id]->width;
    *h = g->as.bitmaps[id]->height;
    return 0;
}
int orc_bitmap_texels(void *gp, int id, int16_t *out) {  // row-major, -1 == None
    Game *g = (Game *)gp;
    if (id < 0 || id >= (int)g->as.bitmaps.size()) return -1;
    auto &b = *g->as.bitmaps[id];
    memcpy(out, b.pixels.data(), b.pixels.size() * sizeof(int16_t));
    return 0;
}
int orc_sky_bitmap_id(void *gp) { return ((Game *)gp)->sky->id; }
int orc_flat_count(void *gp) { return (int)((Game *)gp)->as.flat_list.size(); }
int orc_flat_texels(void *gp, int id, uint8_t *out4096) {
    Game *g = (Game *)gp;
    if (id < 0 || id >= (int)g->as.flat_list.size()) return -1;
    memcpy(out4096, g->as.flat_list[id]->pixels, 4096);
    return 0;
}

// ---- trace of the last orc_render(..., want_trace=1) ----
int orc_trace_count(void *gp) { return (int)((Game *)gp)->trace.size(); }
// ints: kind, phase, bitmap_id|flat_id, light_level, start_x, end_x, offset_x, offset_y, x, clipped_bottom_y, clipped_top_y,
//       bottom_y, top_y, is_sky, height, left, right            (17 ints)
// floats: line.start.x, line.start.y, line.end.x, line.end.y, start_offset, bottom_height, top_height   (7 floats)
int orc_trace_get(void *gp, int i, int32_t *ints17, float *floats7) {
    Game *g = (Game *)gp;
    if (i < 0 || i >= (int)g->trace.size()) return -1;
    const TraceCall &t = g->trace[i];
    int32_t v[17] = {t.kind, t.phase, t.kind == 0 ? t.bitmap_id : t.flat_id, t.light_level, t.start_x, t.end_x, t.offset_x, t.offset_y,
                     t.x, t.clipped_bottom_y, t.clipped_top_y, t.bottom_y, t.top_y, t.is_sky, t.height, t.left, t.right};
    memcpy(ints17, v, sizeof v);
    float f[7] = {t.cl.line.start.x, t.cl.line.start.y, t.cl.line.end.x, t.cl.line.end.y, t.cl.start_offset, t.bottom_height, t.top_height};
    memcpy(floats7, f, sizeof f);
    return 0;
}
int orc_trace_visplane_arrays(void *gp, int i, int16_t *top, int16_t *bottom) {  // W entries each
    Game *g = (Game *)gp;
    if (i < 0 || i >= (int)g->trace.size() || g->trace[i].kind != 1) return -1;
    memcpy(top, g->trace[i].top.data(), g->trace[i].top.size() * 2);
    memcpy(bottom, g->trace[i].bottom.data(), g->trace[i].bottom.size() * 2);
    return 0;
}

// ---- leaf functions on caller-supplied data (property / fuzz tests of the kernels) ----
void orc_diminish_color(const uint8_t *rgb_in, int light_level, int distance, uint8_t *rgb_out) {
    Color c = diminish_color({rgb_in[0], rgb_in[1], rgb_in[2]}, (int16_t)light_level, (int16_t)distance);
    rgb_out[0] = c.r;
    rgb_out[1] = c.g;
    rgb_out[2] = c.b;
}

struct OrcLeafCtx {
    int W, H;
    Assets as;  // only palette used
};
void *orc_leaf_new(int W, int H, const uint8_t *palette768) {
    auto *c = new OrcLeafCtx();
    c->W = W;
    c->H = H;
    for (int i = 0; i < 256; i++) c->as.palette[i] = {palette768[i * 3], palette768[i * 3 + 1], palette768[i * 3 + 2]};
    return c;
}
void orc_leaf_free(void *c) { delete (OrcLeafCtx *)c; }

// render_vertical_bitmap_line on raw arguments. texels row-major int16 (-1 None).
// ints: w, h, light_level, start_x, end_x, offset_x, offset_y, x, clipped_bottom_y, clipped_top_y, bottom_y, top_y (12)
// floats: line.start.x, .y, line.end.x, .y, start_offset, bottom_height, top_height (7)
int orc_leaf_column(void *cp, uint8_t *pixels, const int16_t *texels, const int32_t *i, const float *f) {
    OrcLeafCtx *c = (OrcLeafCtx *)cp;
    return guarded([&] {
        Frame fr(c->W, c->H);
        fr.pixels = pixels;
        Bitmap bm;
        bm.width = (int16_t)i[0];
        bm.height = (int16_t)i[1];
        bm.pixels.assign(texels, texels + (size_t)i[0] * i[1]);
        ClippedLine cl{{{f[0], f[1]}, {f[2], f[3]}}, f[4]};
        render_vertical_bitmap_line(fr, c->as, bm, (int16_t)i[2], cl, i[3], i[4], f[5], f[6], (int16_t)i[5], (int16_t)i[6], i[7], i[8],
                                    i[9], i[10], i[11]);
    });
}
// draw_visplane / draw_sky on raw arguments. top/bottom: W entries. flat: 4096 bytes (ignored for sky);
// sky texels: 256x128 row-major int16 (ignored for flats).
// ints: is_sky, height, light_level, left, right (5); floats: pos.x, pos.y, floor_height, angle (4)
int orc_leaf_visplane(void *cp, uint8_t *pixels, const uint8_t *flat, const int16_t *sky_texels, const int16_t *top, const int16_t *bottom,
                      const int32_t *i, const float *f) {
    OrcLeafCtx *c = (OrcLeafCtx *)cp;
    return guarded([&] {
        Frame fr(c->W, c->H);
        fr.pixels = pixels;
        auto fl = std::make_shared<Flat>();
        fl->name = i[0] ? "F_SKY1" : "FLAT";
        if (flat) memcpy(fl->pixels, flat, 4096);
        Bitmap sky;
        sky.width = 256;
        sky.height = 128;
        if (sky_texels)
            sky.pixels.assign(sky_texels, sky_texels + 256 * 128);
        else
            sky.pixels.assign(256 * 128, -1);
        Visplane vp(fl, (int16_t)i[1], (int16_t)i[2], c->W);
        vp.left = (int16_t)i[3];
        vp.right = (int16_t)i[4];
        vp.top.assign(top, top + c->W);
        vp.bottom.assign(bottom, bottom + c->W);
        Player p{{f[0], f[1]}, f[2], f[3]};
        draw_visplane(fr, c->as, p, sky, vp);
    });
}

}  // extern "C"
