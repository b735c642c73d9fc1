/* drr.h -- C ABI of libdrr.so: the B200 (sm_100a) replacement for the per-pixel back-end of
 * freewilll/doom-rust-renderer.
 *
 * The reference has no FFI of its own; its renderer API is the Rust call
 *     Renderer::new(&mut Pixels, &Map, &MapObjects, &mut Textures, &mut Sprites, sky, &mut Flats, &Palette, &Player, timestamp).render()
 * (src/renderer/mod.rs:37-59,118-136; only call site src/game.rs:505-519) whose one observable effect is filling
 * Pixels.pixels (RGB24, src/renderer/pixels.rs:5-30).  The seam this library plugs into is the set of three leaf call
 * sites of the per-pixel drawers:
 *     src/renderer/segs.rs:234           -> render_vertical_bitmap_line   (solid / lower / upper wall columns)
 *     src/renderer/bitmap_render.rs:109  -> render_vertical_bitmap_line   (masked mid-textures and sprites)
 *     src/renderer/mod.rs:108            -> draw_visplane / draw_sky      (floors, ceilings, sky)
 * At those sites the host, instead of drawing, appends to a per-frame list (drr_emit_*); the list order is the draw
 * order and "last writer wins, transparent texels do not write" is preserved per pixel.  INTEGRATION.md shows the Rust
 * `extern "C"` block and the three one-line call-site changes.
 *
 * Conventions: every function returns 0 on success or a negative DRR_E_* code (never aborts, never unwinds);
 * drr_last_error(ctx) gives a message.  The caller owns every input buffer, the library copies before returning;
 * output buffers are caller-allocated.  One context = one caller thread and one GPU (the reference is single-threaded,
 * its types are !Send); distinct contexts may be driven from distinct threads / processes, one per GPU.
 * No CPU fallback exists: without a CUDA device drr_ctx_create fails with DRR_E_CUDA.
 */
#ifndef DRR_H
#define DRR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRR_OK 0
#define DRR_E_INVALID (-1)   /* bad argument */
#define DRR_E_STATE (-2)     /* call out of sequence (e.g. emit outside frame_begin/frame_end) */
#define DRR_E_CUDA (-3)      /* CUDA runtime error (message in drr_last_error) */
#define DRR_E_NOMEM (-4)     /* host or device allocation failed */
#define DRR_E_ASSET (-5)     /* unknown bitmap / flat id, sky not set, palette missing */
#define DRR_E_IO (-6)        /* scene loader: file / WAD format problem */
#define DRR_E_PANIC (-7)     /* scene front-end: the reference would have panicked on this input */

typedef struct drr_ctx drr_ctx;

/* Player (src/game.rs:41-45) plus cos/sin of `angle` evaluated by the HOST libm (the same sinf/cosf the reference's
 * Vertex::rotate calls, src/map/vertexes.rs:20-25), so the device never evaluates a transcendental. 24 bytes. */
typedef struct {
    float pos_x, pos_y;
    float floor_height;
    float angle;
    float cos_angle, sin_angle;
} drr_view;

#define DRR_PHASE_WALL 0   /* drawn immediately, src/renderer/segs.rs:234 */
#define DRR_PHASE_MASKED 2 /* deferred masked mid-texture or sprite, src/renderer/bitmap_render.rs:109 */

/* Per-seg arguments of render_vertical_bitmap_line (src/renderer/bitmap_render.rs:213-224); the same fields the
 * reference keeps in BitmapRender (bitmap_render.rs:29-45). 48 bytes. */
typedef struct {
    int32_t bitmap_id;   /* handle given to drr_upload_bitmap */
    int16_t light_level; /* sector light level */
    int16_t phase;       /* DRR_PHASE_* (informational; list order is what defines the result) */
    float line_start_x, line_start_y, line_end_x, line_end_y; /* ClippedLine.line, view space */
    float start_offset;                                        /* ClippedLine.start_offset */
    int32_t start_x, end_x;                                    /* screen x of the clipped line ends */
    float bottom_height, top_height;
    int16_t offset_x, offset_y;
} drr_seg_hdr;

/* Per-column arguments (bitmap_render.rs:225-229) == BitmapColumn (bitmap_render.rs:19-25); the reference widens
 * i16 to i32 when it stores them (add_column, :84-99), so i16 is lossless. 10 bytes. */
typedef struct {
    int16_t x;
    int16_t clipped_top_y, clipped_bottom_y;
    int16_t bottom_y, top_y;
} drr_col;

#define DRR_FLAT_SKY (-1)

/* Visplane (src/renderer/visplanes.rs:17-26) minus its arrays. 12 bytes. */
typedef struct {
    int16_t flat_id; /* handle given to drr_upload_flat, or DRR_FLAT_SKY when the flat's name contains "SKY" (visplanes.rs:89) */
    int16_t height;
    int16_t light_level;
    int16_t left, right;
    int16_t reserved;
} drr_visplane_hdr;

/* ---- context ------------------------------------------------------------------------------------------------ */
/* width/height play the role of SCREEN_WIDTH/SCREEN_HEIGHT (src/game.rs:28-29); the f32 projection constants of
 * src/renderer/constants.rs:7-17 are derived from them exactly as the reference does.  max_views = number of
 * framebuffers kept resident in HBM (width*height*3 bytes each). */
int drr_ctx_create(int width, int height, int device_ordinal, int max_views, drr_ctx **out);
void drr_ctx_destroy(drr_ctx *ctx);
const char *drr_last_error(const drr_ctx *ctx); /* ctx may be NULL: last creation error of this thread */
const char *drr_error_name(int code);
/* Use the caller's CUDA stream (a cudaStream_t passed as void*) for all copies and launches; NULL = own stream. */
int drr_set_stream(drr_ctx *ctx, void *cuda_stream);
void *drr_get_stream(drr_ctx *ctx);
/* Tuning / diagnostic knobs (nothing the results depend on).  Each knob has an environment variable that is read ONCE, by
 * drr_ctx_create -- never on the launch path --, and drr_set_knob changes it afterwards; value 0 = built-in default.
 *   "submit_chunks" (DRR_SUBMIT_CHUNKS), "submit_one_stream" (DRR_SUBMIT_ONE_STREAM), "submit_trace" (DRR_SUBMIT_TRACE),
 *   "fe_trace" (DRR_FE_TRACE), "fe_two_pass" (DRR_FE_TWO_PASS), "fe_slab_div" (DRR_FE_SLAB_DIV), "fe_cap_renders",
 *   "fe_cap_dsegs", "fe_cap_allcols_per_w" (DRR_FE_CAP_*), "dbg" (DRR_DBG; -DDRR_DBG_KNOBS builds only). */
int drr_set_knob(drr_ctx *ctx, const char *name, int value);

/* ---- assets (uploaded once, device resident) ------------------------------------------------------------------ */
int drr_upload_palette(drr_ctx *ctx, const uint8_t rgb[768]);                                   /* Palette::new, src/graphics/palette.rs:11-28 */
int drr_upload_bitmap(drr_ctx *ctx, int id, int w, int h, const int16_t *texels_rowmajor);       /* Bitmap, src/graphics/bitmap.rs:11-15; -1 == None */
int drr_upload_flat(drr_ctx *ctx, int id, const uint8_t px[4096]);                               /* Flat, src/graphics/flats.rs:19-22 */
int drr_set_sky(drr_ctx *ctx, int bitmap_id);                                                    /* sky_texture, src/renderer/mod.rs:32 (must be 256x128) */

/* ---- recording one frame == one Renderer::render() ------------------------------------------------------------ */
int drr_reset(drr_ctx *ctx); /* forget all recorded frames */
int drr_frame_begin(drr_ctx *ctx, int view_idx, const drr_view *view);
int drr_emit_columns(drr_ctx *ctx, const drr_seg_hdr *hdr, const drr_col *cols, int n);          /* replaces n calls of render_vertical_bitmap_line */
int drr_emit_visplane(drr_ctx *ctx, const drr_visplane_hdr *hdr, const int16_t *top, const int16_t *bottom);
/* top/bottom point at the entries for x = hdr->left .. hdr->right (right-left+1 values each), passed through
 * UNCLAMPED exactly as the reference stores them (quirk Q3, visplanes.rs:36-37,60-64). */
int drr_frame_end(drr_ctx *ctx);
int drr_frame_abort(drr_ctx *ctx); /* discard the frame being recorded (the view index becomes free again) */

/* ---- recorders: the same recording off-context ------------------------------------------------------------------ */
/* A recorder records frames exactly like drr_frame_begin .. drr_frame_end but into private memory, so several threads can
 * run the front-end at once (one recorder per thread; the context's asset tables are only read while recording, so assets
 * must be uploaded before).  drr_append, called from the context's own thread, moves the recorder's frames to the end of
 * the context's lists and empties the recorder; view indices must still be unique per drr_reset. */
typedef struct drr_recorder drr_recorder;
int drr_recorder_create(drr_ctx *ctx, drr_recorder **out);
void drr_recorder_destroy(drr_recorder *rec);
const char *drr_recorder_last_error(const drr_recorder *rec);
int drr_recorder_frame_begin(drr_recorder *rec, int view_idx, const drr_view *view);
int drr_recorder_emit_columns(drr_recorder *rec, const drr_seg_hdr *hdr, const drr_col *cols, int n);
int drr_recorder_emit_visplane(drr_recorder *rec, const drr_visplane_hdr *hdr, const int16_t *top, const int16_t *bottom);
int drr_recorder_frame_end(drr_recorder *rec);
int drr_recorder_frame_abort(drr_recorder *rec);
int drr_append(drr_ctx *ctx, drr_recorder *rec);

/* ---- execution ---------------------------------------------------------------------------------------------- */
/* drr_upload_lists and drr_submit return while their copies out of the context's pinned host lists may still be running; the
 * library waits for them itself before the lists are touched again (drr_reset, drr_frame_begin, drr_append), so recording
 * the next batch right after a submit is safe. */
int drr_upload_lists(drr_ctx *ctx); /* async H2D of everything recorded since drr_reset (from pinned staging) */
int drr_draw(drr_ctx *ctx);         /* async: render every uploaded frame into its framebuffer (+ per-frame checksum) */
int drr_submit(drr_ctx *ctx);       /* drr_upload_lists + drr_draw */
int drr_sync(drr_ctx *ctx);
int drr_read_framebuffer(drr_ctx *ctx, int view_idx, uint8_t *out_rgb24);    /* width*height*3 bytes, row-major RGB24 == Pixels.pixels */
int drr_read_checksums(drr_ctx *ctx, int first_view, int count, uint64_t *out);
/* Per-frame checksum.  The frame's little-endian u32 words (zero-padded) are taken in groups of 12 (48 bytes = 16 pixels):
 *   s_g = sum_j w[12g + j] * ((2j + 1) * C) mod 2^32,   checksum = sum_g s_g * ((g + 1) * C mod 2^32) mod 2^64,   C = 0x9E3779B1
 * (every multiplier is odd: a change of any single word changes the sum).  drr_checksum_host computes the same on a host
 * buffer. */
uint64_t drr_checksum_host(const uint8_t *rgb24, uint64_t nbytes);
/* Export side (the reference presents Pixels.pixels through SDL, src/game.rs:500-533): the framebuffer
 * drr_read_framebuffer returns is what an SDL RGB24 streaming texture takes as is (pitch = width*3), and drr_read_crc32 gives
 * the standard CRC-32 (zlib / PNG polynomial 0xEDB88320, as zlib.crc32 of the frame's width*height*3 bytes) of the resident
 * frames, computed on the device by a separate kernel -- chunk CRCs folded with the usual GF(2) shift -- so frames can be
 * compared with files without copying them back.  (The fused per-frame checksum above stays the cheap one: it rides on the
 * write-out's registers; a CRC cannot, its chunks combine in order.) */
int drr_read_crc32(drr_ctx *ctx, int first_view, int count, uint32_t *out);

/* ---- introspection for benchmarks ----------------------------------------------------------------------------- */
typedef struct {
    uint64_t frames;              /* frames recorded */
    uint64_t seg_headers;         /* drr_emit_columns calls */
    uint64_t column_records;      /* drr_col records */
    uint64_t visplanes;           /* drr_emit_visplane calls */
    uint64_t visplane_columns;    /* sum of right-left+1 */
    uint64_t drawlist_bytes_algorithmic; /* 24*frames + 48*seg_headers + 10*column_records + 12*visplanes + 4*visplane_columns */
    uint64_t device_list_bytes;   /* bytes of the column-binned device representation actually uploaded */
    uint64_t spans;               /* resolved spans (opaque + masked) */
    uint64_t kernel_launches;     /* kernels launched by this context so far */
} drr_stats;
int drr_get_stats(drr_ctx *ctx, drr_stats *out);
/* Time `iters` back-to-back drr_draw() passes with CUDA events on the context's stream; returns average milliseconds of
 * the whole pass and of the two kernels separately (setup_ms, march_ms may be NULL). */
int drr_time_draw(drr_ctx *ctx, int iters, float *total_ms, float *setup_ms, float *march_ms);

/* Per-kernel device times of the drr_draw()/drr_submit() calls made between begin and end (CUDA events on the context's
 * stream around each kernel; at most max_steps draws are profiled).  Totals in milliseconds. */
int drr_profile_begin(drr_ctx *ctx, int max_steps);
int drr_profile_end(drr_ctx *ctx, int *steps, float *setup_ms_total, float *march_ms_total);

/* ---- device front-end: Renderer::render()'s BSP walk, seg clipping, occlusion arrays and --- */
/* visplane building, map-object projection / clipping / ordering for a whole batch of viewpoints ON THE GPU
 * (csrc/drr_frontend.cuh, one warp per viewpoint): the draw lists are written where drr_draw reads them and never cross
 * PCIe.  Covers every phase: walls, visplanes, map objects, masked mid-textures (src/renderer/mod.rs:69-136,
 * segs.rs:121-597, misc.rs:13-161, sidedef_visplanes.rs, map_objects.rs:19-241, bitmap_render.rs:101-188).
 * The map goes up as flat tables (what the loaders under src/map/ read, names resolved to the handles given to
 * drr_upload_bitmap / drr_upload_flat): */
typedef struct { float x, y, dx, dy; int32_t right, left; } drr_fe_node;            /* src/map/nodes.rs; child >= 0 node, < 0 ~subsector */
typedef struct { int32_t first_seg, count; } drr_fe_subsector;                      /* src/map/subsectors.rs */
typedef struct { float v1x, v1y, v2x, v2y; int32_t linedef; int16_t direction, offset; } drr_fe_seg; /* src/map/segs.rs */
typedef struct { int32_t front, back, flags; } drr_fe_linedef;                      /* sidedef indices, -1 = none */
typedef struct { float x_offset, y_offset; int32_t upper, lower, middle, sector; } drr_fe_sidedef;  /* bitmap ids; -1 = "-", -2 = unknown name */
typedef struct {
    int16_t floor_height, ceiling_height, light_level;
    int16_t ceiling_name_has_sky;      /* the sector's ceiling texture NAME contains "SKY" (segs.rs:464-469) */
    int16_t floor_flat, ceiling_flat;  /* flat ids for the batch's timestamp (animation resolved, flats.rs:103-111); -2 = lump missing */
    int16_t floor_is_sky, ceiling_is_sky; /* the resolved flat's name contains "SKY" (visplanes.rs:89) */
} drr_fe_sector;
typedef struct { /* a non-null map object at tic 0 (src/map_objects.rs:25-50, src/info.rs): everything view-independent resolved */
    float x, y, angle;
    int32_t sector;        /* the sector its position lies in (src/renderer/bsp.rs:9-44), -1 = outside the map */
    int32_t full_bright;
    int32_t rotate;        /* its sprite frame has 8 rotations (src/graphics/sprites.rs:26-97) */
    int32_t bitmap[8];     /* bitmap id per rotation (entry 0 when !rotate) */
    int16_t top_offset[8]; /* Picture.top_offset per rotation */
} drr_fe_thing;
typedef struct {
    const drr_fe_node *nodes;           int32_t n_nodes;
    const drr_fe_subsector *subsectors; int32_t n_subsectors;
    const drr_fe_seg *segs;             int32_t n_segs;
    const drr_fe_linedef *linedefs;     int32_t n_linedefs;
    const drr_fe_sidedef *sidedefs;     int32_t n_sidedefs;
    const drr_fe_sector *sectors;       int32_t n_sectors;
    const drr_fe_thing *things;         int32_t n_things; /* may be NULL / 0 */
} drr_fe_map;
int drr_fe_upload_map(drr_ctx *ctx, const drr_fe_map *map); /* assets must be uploaded before (ids are resolved here) */
uint32_t drr_fe_map_id(drr_ctx *ctx); /* 0 = no map; otherwise a number that changes with every drr_fe_upload_map (callers cache on it) */
/* n x Renderer::new(..., player at xya[i], ...).render() on the device, view indices first_view_idx .. +n-1.  Replaces
 * every frame recorded since drr_reset (call drr_reset first); afterwards drr_draw() renders the batch.  status[i]
 * (may be NULL) = DRR_OK or DRR_E_PANIC (the reference would have panicked on that viewpoint: it gets no frame). */
int drr_fe_emit_views(drr_ctx *ctx, int first_view_idx, const float *xya, int n, int phases, int *status);
/* How the last drr_fe_emit_views ran and its device times in ms (CUDA events on the context's stream).  Mode 1 = single
 * pass: every view writes into its own slab, and the draw kernels read the slabs in place (emit_ms = the front-end
 * kernels, count_ms = the gather of the frames' View records; DRR_FE_COMPACT=1 copies the slabs into dense lists first,
 * count_ms is then that copy).  Mode 2 = two passes (a view outgrew its slab, or DRR_FE_TWO_PASS is set): count pass,
 * then emit pass straight into dense lists.  Dense lists hold the same bytes in both modes. */
int drr_fe_last_times(drr_ctx *ctx, float *count_ms, float *emit_ms);
int drr_fe_last_mode(drr_ctx *ctx);

/* ---- host front-end: the reference's Renderer for a WAD map, emitting through the functions above -------------- */
/* Mirrors Game::new's asset/map loading (src/game.rs:118-196) without SDL. */
typedef struct drr_scene drr_scene;
int drr_scene_load(const char *wad_path, const char *map_name, int width, int height, drr_scene **out);
void drr_scene_free(drr_scene *scene);
const char *drr_scene_last_error(const drr_scene *scene); /* scene may be NULL */
int drr_scene_upload_assets(drr_scene *scene, drr_ctx *ctx); /* palette, every bitmap/flat the map can reference, sky */
int drr_scene_player_start(drr_scene *scene, float out_xya[3]); /* Player1Start, src/game.rs:151-157 */
/* The time axis: put the world `tic` game ticks (1/35 s, src/game.rs:32) after the start of the game -- the
 * sector light effects (src/lights.rs, src/thinkers.rs:14-76) and the map objects' state machines (src/map_objects.rs:63-95)
 * stepped tic by tic as Game::tick does (src/game.rs:456-482).  The reference draws its random numbers from
 * rand::thread_rng(); here they come from one PCG32 stream seeded with `seed`, consumed in the reference's thinker order
 * (gen_range(lo..hi) = lo + next % (hi - lo)), so a (tic, seed) pair always gives the same world.  tic 0 = the WAD as loaded.
 * Affects every later drr_scene_emit_* call.  The animated flats follow the `timestamp` argument of those calls. */
int drr_scene_set_tic(drr_scene *scene, uint32_t tic, uint64_t seed);
int drr_scene_counts(drr_scene *scene, int *n_sectors, int *n_objects);
/* sector_lights: n_sectors values; object_states4: n_objects x (sprite index, frame, full_bright, is S_NULL); either may be NULL */
int drr_scene_world_state(drr_scene *scene, int16_t *sector_lights, int32_t *object_states4);
#define DRR_PHASES_WALLS 1
#define DRR_PHASES_PLANES 2
#define DRR_PHASES_MASKED 4
#define DRR_PHASES_ALL 7
/* One Renderer::new(..., player, timestamp).render() for the player at (x, y, angle): frame_begin, emits, frame_end.
 * `phases` gates which leaf call sites emit (config 3's walls-only / flats-only split). */
int drr_scene_emit_view(drr_scene *scene, drr_ctx *ctx, int view_idx, float x, float y, float angle, float timestamp, int phases);
/* The same for n viewpoints (xya = n x (x, y, angle)) with view indices first_view_idx .. first_view_idx+n-1, the front-end
 * running on `nthreads` worker threads (0 = one per host core) through recorders.  status[i] (may be NULL) = DRR_OK or
 * DRR_E_PANIC (nothing recorded for that viewpoint). */
int drr_scene_emit_views(drr_scene *scene, drr_ctx *ctx, int first_view_idx, const float *xya, int n, float timestamp, int phases,
                         int nthreads, int *status);

/* drr_fe_upload_map (this scene's map, flats resolved for `timestamp`) + drr_fe_emit_views. */
int drr_scene_emit_views_device(drr_scene *scene, drr_ctx *ctx, int first_view_idx, const float *xya, int n, float timestamp, int phases,
                                int *status);

#ifdef __cplusplus
}
#endif
#endif /* DRR_H */
