// hp_dataset.cu -- host-side reader of the reference's recorded datasets, feeding the batched
// entry points.  Replaces load_dataset (include/dataset.h:109-163) for the files DepthDataStreamOut writes
// (dataset.h:62-105):
//   <base>.json  DatasetInfo: camera intrinsics and header fields (dataset.h:21-37, misc_image.h:57)
//   <base>.rs    headerless 16-bit depth, width x height per frame (plus width x height IR bytes per frame when the
//                deprecated "hasir" interleave is set, dataset.h:135)
//   <base>.ir    optional, 8-bit IR, width x height per frame
//   <base>.pose  optional, ASCII: per frame pose_array_size x (position xyz, orientation xyzw)
// The reference copies every frame into std::vectors one istream::read at a time; here the binary files are mapped and
// frames are copied in one memcpy per request straight into the caller's (pinned) batch buffer, or -- for datasets
// already reduced to 64x64 crops (train-cnn.cpp:31-34) -- handed to hp_eval_depth_batch as they lie in the page cache.
// No GPU code in this file; it is compiled by nvcc with the rest of the library so that there is one .so.
#include "../../include/handposedd.h"
#include "hp_common.cuh"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

using hp::set_error;

namespace {

struct Mapped {
    const uint8_t *p = nullptr;
    size_t n = 0;
    bool open = false;    // the file exists (an empty file is open with n == 0)
    void map(const std::string &path)
    {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        open = true;
        struct stat st;
        if (fstat(fd, &st) == 0 && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                p = (const uint8_t *)m;
                n = (size_t)st.st_size;
            }
        }
        ::close(fd);
    }
    void unmap()
    {
        if (p) munmap((void *)p, n);
        p = nullptr;
        n = 0;
    }
};

// ---- the subset of JSON that DatasetInfo needs: one object of numbers, strings, booleans, arrays and objects --------
struct JsonCursor {
    const char *s, *e;
    void ws() { while (s < e && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r')) s++; }
    bool lit(char c) { ws(); if (s < e && *s == c) { s++; return true; } return false; }
};

struct JsonValue {
    enum Kind { NUL, NUM, STR, TRUE_, FALSE_, ARR, OBJ } kind = NUL;
    std::string text;                                       // NUM: the token as written; STR: the unescaped contents
    std::vector<JsonValue> items;                           // ARR
    std::vector<std::pair<std::string, JsonValue>> fields;  // OBJ
    const JsonValue &at(size_t i) const { static const JsonValue null; return (kind == ARR && i < items.size()) ? items[i] : null; }
    const JsonValue &at(const char *key) const
    {
        static const JsonValue null;
        if (kind == OBJ)
            for (auto &f : fields)
                if (f.first == key) return f.second;
        return null;
    }
    // json.h:104-107: a number is converted from its text by stream extraction, anything else yields T()
    float f32() const { return kind == NUM ? strtof(text.c_str(), nullptr) : 0.f; }
    int i32() const { return kind == NUM ? (int)strtol(text.c_str(), nullptr, 10) : 0; }
};

static bool parse_value(JsonCursor &c, JsonValue &v, int depth);

static bool parse_string(JsonCursor &c, std::string &out)
{
    if (!c.lit('"')) return false;
    while (c.s < c.e && *c.s != '"') {
        if (*c.s == '\\' && c.s + 1 < c.e) {
            c.s++;
            switch (*c.s) {
            case 'n': out.push_back('\n'); break;
            case 't': out.push_back('\t'); break;
            case 'r': out.push_back('\r'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            default: out.push_back(*c.s); break;   // \" \\ \/ (and \uXXXX kept verbatim: not used by DatasetInfo)
            }
            c.s++;
        } else
            out.push_back(*c.s++);
    }
    return c.lit('"');
}

static bool parse_value(JsonCursor &c, JsonValue &v, int depth)
{
    if (depth > 16) return false;
    c.ws();
    if (c.s >= c.e) return false;
    const char ch = *c.s;
    if (ch == '{') {
        c.s++;
        v.kind = JsonValue::OBJ;
        if (c.lit('}')) return true;
        do {
            std::string key;
            JsonValue item;
            if (!parse_string(c, key) || !c.lit(':') || !parse_value(c, item, depth + 1)) return false;
            v.fields.emplace_back(std::move(key), std::move(item));
        } while (c.lit(','));
        return c.lit('}');
    }
    if (ch == '[') {
        c.s++;
        v.kind = JsonValue::ARR;
        if (c.lit(']')) return true;
        do {
            JsonValue item;
            if (!parse_value(c, item, depth + 1)) return false;
            v.items.push_back(std::move(item));
        } while (c.lit(','));
        return c.lit(']');
    }
    if (ch == '"') {
        v.kind = JsonValue::STR;
        return parse_string(c, v.text);
    }
    auto word = [&](const char *w, JsonValue::Kind k) {
        const size_t n = strlen(w);
        if ((size_t)(c.e - c.s) >= n && memcmp(c.s, w, n) == 0) { c.s += n; v.kind = k; return true; }
        return false;
    };
    if (word("true", JsonValue::TRUE_) || word("false", JsonValue::FALSE_) || word("null", JsonValue::NUL)) return true;
    const char *b = c.s;
    while (c.s < c.e && (strchr("+-.eE", *c.s) || (*c.s >= '0' && *c.s <= '9'))) c.s++;
    if (c.s == b) return false;
    v.kind = JsonValue::NUM;
    v.text.assign(b, c.s);
    return true;
}

}  // namespace

struct hp_dataset {
    hp_dataset_info info;
    Mapped rs, ir;
    std::vector<float> poses;   // [n_frames][pose_array_size][7]
    size_t pixels = 0, rs_stride = 0;
};

// `in >> float` over the whole .pose text in file order (geometric.h:133,139), with the stream's failure behaviour:
// at end of file nothing more is stored; at a token that is not a number the element being read becomes 0 (C++11
// num_get) -- and in both cases everything after it keeps the value a default-constructed Pose has: position (0,0,0),
// orientation (0,0,0,1).
static void parse_poses(const Mapped &m, int64_t n_frames, int np, std::vector<float> &out)
{
    out.assign((size_t)n_frames * np * 7, 0.f);
    for (size_t i = 6; i < out.size(); i += 7) out[i] = 1.f;
    if (!m.open || out.empty()) return;
    std::string text((const char *)m.p, m.n);   // NUL-terminated copy for strtof
    const char *s = text.c_str();
    for (size_t i = 0; i < out.size(); i++) {
        while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\v' || *s == '\f') s++;
        if (*s == 0) return;   // end of file: the stream's sentry fails before num_get runs, the element keeps its default
        // num_get only accumulates characters a decimal float can contain, so "nan", "inf" and hex floats fail (or stop
        // early) exactly as they do for the reference's `in >> v[i]`
        char tok[64];
        size_t len = 0;
        while (len + 1 < sizeof(tok) && s[len] && (strchr("+-.eE", s[len]) || (s[len] >= '0' && s[len] <= '9'))) { tok[len] = s[len]; len++; }
        tok[len] = 0;
        char *tend = nullptr;
        const float v = strtof(tok, &tend);
        if (tend == tok) {   // a token that is not a number: failbit, zero stored
            out[i] = 0.f;
            return;
        }
        const char *end = s + (tend - tok);
        out[i] = v;
        s = end;
    }
}

extern "C" {

int hp_dataset_open(const char *basename, int pose_array_size, hp_dataset **out)
{
    if (!basename || !out || pose_array_size < 0) { set_error("bad argument"); return HP_ERR_INVALID; }
    *out = nullptr;
    const std::string base(basename);
    hp_dataset *d = new hp_dataset;
    memset(&d->info, 0, sizeof(d->info));
    d->rs.map(base + ".rs");
    if (!d->rs.open) {   // dataset.h:114-115
        set_error("unable to open %s.rs", basename);
        delete d;
        return HP_ERR_IO;
    }
    Mapped js;
    js.map(base + ".json");
    if (!js.open) {      // dataset.h:117-118
        set_error("%s.json not found", basename);
        d->rs.unmap();
        delete d;
        return HP_ERR_IO;
    }
    JsonValue root;
    JsonCursor cur{(const char *)js.p, (const char *)js.p + js.n};
    const bool ok = js.n > 0 && parse_value(cur, root, 0) && root.kind == JsonValue::OBJ;
    js.unmap();
    if (!ok) {
        set_error("%s.json is not a JSON object", basename);
        d->rs.unmap();
        delete d;
        return HP_ERR_IO;
    }
    hp_dataset_info &I = d->info;
    const JsonValue &cam = root.at("dcamera");            // visit_fields(DCamera), misc_image.h:57
    I.width = cam.at("dims").at((size_t)0).i32();
    I.height = cam.at("dims").at(1).i32();
    for (int k = 0; k < 2; k++) {
        I.focal[k] = cam.at("focal").at(k).f32();
        I.principal[k] = cam.at("principal").at(k).f32();
        I.rgb_dim[k] = root.at("rgb_dim").at(k).i32();     // visit_fields(DatasetInfo), dataset.h:32-37
        I.feye_dim[k] = root.at("feyedim").at(k).i32();
    }
    I.depth_scale = cam.at("depth_scale").f32();
    for (int k = 0; k < 4; k++) I.mplane[k] = root.at("mplane").at(k).f32();
    I.hasir = root.at("hasir").kind == JsonValue::TRUE_;
    I.segment_scale = root.at("segment_scale").f32();
    snprintf(I.camtype, sizeof(I.camtype), "%s", root.at("camtype").kind == JsonValue::STR ? root.at("camtype").text.c_str() : "");
    if (I.width <= 0 || I.height <= 0) {
        // the reference would spin forever here (zero-byte reads never reach end of file, dataset.h:129-133)
        set_error("%s.json: dcamera.dims is %d x %d", basename, I.width, I.height);
        d->rs.unmap();
        delete d;
        return HP_ERR_IO;
    }
    d->pixels = (size_t)I.width * I.height;
    d->rs_stride = d->pixels * 2 + (I.hasir ? d->pixels : 0);   // dataset.h:133-136: a trailing partial frame is dropped
    I.n_frames = (int64_t)(d->rs.n / d->rs_stride);
    I.pose_array_size = pose_array_size;
    d->ir.map(base + ".ir");
    I.has_ir_file = d->ir.open;
    Mapped pose;
    pose.map(base + ".pose");
    I.has_pose_file = pose.open;
    parse_poses(pose, I.n_frames, pose_array_size, d->poses);
    pose.unmap();
    *out = d;
    return HP_OK;
}

int hp_dataset_get_info(const hp_dataset *ds, hp_dataset_info *info)
{
    if (!ds || !info) { set_error("bad argument"); return HP_ERR_INVALID; }
    *info = ds->info;
    return HP_OK;
}

int hp_dataset_read(hp_dataset *ds, int64_t first, int64_t count, uint16_t *depth, uint8_t *ir, float *poses)
{
    if (!ds || first < 0 || count < 0 || first + count > ds->info.n_frames) { set_error("frame range out of bounds"); return HP_ERR_INVALID; }
    const size_t px = ds->pixels;
    for (int64_t i = 0; i < count; i++) {
        const uint8_t *src = ds->rs.p + (size_t)(first + i) * ds->rs_stride;
        if (depth) {
            if (!ds->info.hasir && i == 0) {   // frames are contiguous in the file: one copy for the whole request
                memcpy(depth, src, (size_t)count * px * 2);
            } else if (ds->info.hasir)
                memcpy(depth + (size_t)i * px, src, px * 2);
        }
        if (ir) {
            uint8_t *dst = ir + (size_t)i * px;
            memset(dst, 0, px);                                        // dataset.h:132: zero-initialised
            if (ds->info.hasir) memcpy(dst, src + px * 2, px);         // dataset.h:135
            if (ds->ir.open) {                                         // dataset.h:137-138: a short .ir file fills a prefix
                const size_t off = (size_t)(first + i) * px;
                if (off < ds->ir.n) memcpy(dst, ds->ir.p + off, ds->ir.n - off < px ? ds->ir.n - off : px);
            }
        }
    }
    if (poses && ds->info.pose_array_size > 0)
        memcpy(poses, ds->poses.data() + (size_t)first * ds->info.pose_array_size * 7, (size_t)count * ds->info.pose_array_size * 7 * sizeof(float));
    return HP_OK;
}

int hp_dataset_eval_depth(hp_net *net, hp_dataset *ds, int64_t first, int64_t count, float dmin, float dmax, float *y, float *decoded,
                          int precision)
{
    if (!net || !ds || first < 0 || count < 0 || first + count > ds->info.n_frames) { set_error("bad argument"); return HP_ERR_INVALID; }
    if (ds->info.width != 64 || ds->info.height != 64) {
        set_error("frames are %d x %d: only datasets already reduced to 64x64 hand crops (train-cnn.cpp:31-34) can be fed to the net directly; "
                  "segmentation (HandSegmentVR) stays on the host", ds->info.width, ds->info.height);
        return HP_ERR_UNSUPPORTED;
    }
    if (count == 0) return HP_OK;
    if (!ds->info.hasir)   // frames lie back to back in the mapping
        return hp_eval_depth_batch(net, (const uint16_t *)(ds->rs.p + (size_t)first * ds->rs_stride), count, ds->info.depth_scale, dmin, dmax, y, decoded,
                                   precision);
    std::vector<uint16_t> buf((size_t)count * 4096);
    if (int rc = hp_dataset_read(ds, first, count, buf.data(), nullptr, nullptr)) return rc;
    return hp_eval_depth_batch(net, buf.data(), count, ds->info.depth_scale, dmin, dmax, y, decoded, precision);
}

void hp_dataset_close(hp_dataset *ds)
{
    if (!ds) return;
    ds->rs.unmap();
    ds->ir.unmap();
    delete ds;
}

}  // extern "C"
