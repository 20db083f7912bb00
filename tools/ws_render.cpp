// ws_render -- offline dataset renderer on top of the C ABI, in C++ (the compiled-language counterpart of the reference's
// `render` binary, bin/render.rs:14-180; scripts/render_scene.py is the same tool in Python).
//
//   ws_render <input.ply|.npz> <cameras.json> <img_out> [--max-sh-deg N]   renders the test split, then the train split, to PNG
//   ws_render --parse-scene <cameras.json>                              prints what Scene::from_json + Into<PerspectiveCamera> give (no GPU)
//   ws_render --png-selftest <out.png> <width> <height>                 writes a synthetic f16 frame through the pixel conversion + PNG writer (no GPU)
//
//   ws_render --parse-npz <file.npz>                                    lists the members of a .npz (name, dtype, shape, FNV-1a of the bytes; no GPU)
//
// Input: .ply, or the compressed .npz of io/npz.rs (tools/npz_reader.hpp decodes the zip + npy container -- deflated members
// need the zlib build --, the arrays go to the GPU through ws_pointcloud_create_from_c3dgs).  PNG files use stored
// (uncompressed) deflate blocks, so writing them needs no zlib.
#include "../include/websplat_b200.hpp"
#include "npz_reader.hpp"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <sys/stat.h>

namespace {

// ---- a small JSON reader: enough for cameras.json (array of flat objects with numbers, strings and nested number arrays)
struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    double num = 0; bool b = false; std::string str;
    std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;
    const Json &at(const std::string &key) const
    {
        for (auto &kv : obj) if (kv.first == key) return kv.second;
        throw std::runtime_error("cameras.json: missing field '" + key + "'");
    }
};

class JsonParser {
public:
    explicit JsonParser(const std::string &s) : s_(s) {}
    Json parse() { Json v = value(); ws(); if (p_ != s_.size()) fail("trailing characters"); return v; }

private:
    const std::string &s_; size_t p_ = 0;
    [[noreturn]] void fail(const char *what) const { throw std::runtime_error("cameras.json: " + std::string(what) + " at byte " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) p_++; }
    bool eat(char c) { ws(); if (p_ < s_.size() && s_[p_] == c) { p_++; return true; } return false; }
    Json value()
    {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        const char c = s_[p_];
        Json v;
        if (c == '{') {
            p_++; v.kind = Json::Object;
            if (eat('}')) return v;
            do { ws(); Json k = string_(); if (!eat(':')) fail("expected ':'"); v.obj.emplace_back(k.str, value()); } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            p_++; v.kind = Json::Array;
            if (eat(']')) return v;
            do { v.arr.push_back(value()); } while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            v = string_();
        } else if (s_.compare(p_, 4, "true") == 0) { p_ += 4; v.kind = Json::Bool; v.b = true; }
        else if (s_.compare(p_, 5, "false") == 0) { p_ += 5; v.kind = Json::Bool; }
        else if (s_.compare(p_, 4, "null") == 0) { p_ += 4; }
        else {
            char *end = nullptr;
            v.num = std::strtod(s_.c_str() + p_, &end);
            if (end == s_.c_str() + p_) fail("unexpected character");
            p_ = (size_t)(end - s_.c_str()); v.kind = Json::Number;
        }
        return v;
    }
    Json string_()
    {
        if (p_ >= s_.size() || s_[p_] != '"') fail("expected string");
        p_++;
        Json v; v.kind = Json::String;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\' && p_ < s_.size()) {
                const char e = s_[p_++];
                switch (e) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break; case 'b': c = '\b'; break; case 'f': c = '\f'; break;
                             case 'u': p_ += 4; c = '?'; break; default: c = e; }
            }
            v.str.push_back(c);
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return v;
    }
};

std::string read_file(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    std::ostringstream ss; ss << f.rdbuf();
    return ss.str();
}

// serde: SceneCamera { id, img_name, width, height, position [3], rotation [[3];3], fx, fy } (scene.rs:11-24)
ws::Scene scene_from_json(const std::string &text)
{
    const Json root = JsonParser(text).parse();
    if (root.kind != Json::Array) throw std::runtime_error("cameras.json: expected an array of cameras");
    std::vector<ws::SceneCamera> cams;
    for (const Json &e : root.arr) {
        ws::SceneCamera c;
        c.id = (size_t)e.at("id").num; c.img_name = e.at("img_name").str;
        c.width = (uint32_t)e.at("width").num; c.height = (uint32_t)e.at("height").num;
        const Json &p = e.at("position"), &r = e.at("rotation");
        if (p.arr.size() != 3 || r.arr.size() != 3) throw std::runtime_error("cameras.json: position / rotation must have 3 entries");
        for (int i = 0; i < 3; i++) {
            c.position[i] = (float)p.arr[i].num;
            if (r.arr[i].arr.size() != 3) throw std::runtime_error("cameras.json: rotation rows must have 3 entries");
            for (int j = 0; j < 3; j++) c.rotation[i][j] = (float)r.arr[i].arr[j].num;
        }
        c.fx = (float)e.at("fx").num; c.fy = (float)e.at("fy").num;
        cams.push_back(c);
    }
    return ws::Scene::from_file_order(std::move(cams));
}

// ---- compressed .npz -> ws_c3dgs_arrays: NpzReader::new + read (io/npz.rs:29-56,58-160), minus the GPU part ---------------
const npz::Array &need(const std::map<std::string, npz::Array> &m, const char *name, const char *descr)
{
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error(std::string("array '") + name + "' missing");                  // io/npz.rs:265-275
    if (it->second.descr != descr) throw std::runtime_error(std::string("array '") + name + "' has dtype " + it->second.descr + ", expected " + descr);
    return it->second;
}
const npz::Array *maybe(const std::map<std::string, npz::Array> &m, const char *name)
{
    auto it = m.find(name);
    return it == m.end() ? nullptr : &it->second;
}
double scalar_or(const std::map<std::string, npz::Array> &m, const std::string &name, double dflt)
{
    auto it = m.find(name);
    return (it == m.end() || it->second.count() == 0) ? dflt : it->second.as_double(0);
}

// the descriptor points into `m` (and into gi / fi when an index array had to be widened): keep them alive
ws_c3dgs_arrays c3dgs_from_members(const std::map<std::string, npz::Array> &m, std::vector<int32_t> &gi, std::vector<int32_t> &fi)
{
    ws_c3dgs_arrays d;
    std::memset(&d, 0, sizeof d);
    const npz::Array &xyz = need(m, "xyz", "<f2"), &opacity = need(m, "opacity", "|i1");
    const npz::Array &scaling = need(m, "scaling", "|i1"), &rotation = need(m, "rotation", "|i1");
    const npz::Array &dc = need(m, "features_dc", "|i1"), &rest = need(m, "features_rest", "|i1");
    const size_t n = xyz.count() / 3;
    if (opacity.count() != n) throw std::runtime_error("opacity has a different length than xyz");
    d.xyz = xyz.data.data(); d.opacity = reinterpret_cast<const int8_t *>(opacity.data.data()); d.num_points = n;
    const bool has_sf = maybe(m, "scaling_factor_scale") != nullptr;                                           // io/npz.rs:88-96
    if (has_sf) {
        const npz::Array &sf = need(m, "scaling_factor", "|i1");
        if (sf.count() != n) throw std::runtime_error("scaling_factor has a different length than xyz");
        d.scaling_factor = reinterpret_cast<const int8_t *>(sf.data.data());
    }
    auto indices = [&](const char *name, std::vector<int32_t> &store) -> const int32_t * {
        const npz::Array *a = maybe(m, name);
        if (!a) return nullptr;
        if (a->count() != n) throw std::runtime_error(std::string(name) + " has a different length than xyz");
        if (a->descr == "<i4") return reinterpret_cast<const int32_t *>(a->data.data());
        store.resize(n);
        for (size_t i = 0; i < n; i++) store[i] = (int32_t)a->as_double(i);
        return store.data();
    };
    d.gaussian_indices = indices("gaussian_indices", gi);
    d.feature_indices = indices("feature_indices", fi);
    if (scaling.count() / 3 != rotation.count() / 4) throw std::runtime_error("scaling / rotation lengths differ");
    d.scaling = reinterpret_cast<const int8_t *>(scaling.data.data()); d.rotation = reinterpret_cast<const int8_t *>(rotation.data.data());
    d.num_covars = rotation.count() / 4;
    // sh degree from features_rest.shape[1] + 1 (io/npz.rs:33-37)
    const size_t ncoef = rest.shape.size() >= 2 ? rest.shape[1] + 1 : 1;
    size_t root = 0; while (root * root < ncoef) root++;
    if (root * root != ncoef || root == 0 || root > 4) throw std::runtime_error("num sh coefs not valid");
    d.sh_deg = (uint32_t)root - 1;
    d.num_features = dc.count() / 3;
    if (rest.count() != d.num_features * (ncoef - 1) * 3) throw std::runtime_error("features_rest / features_dc lengths differ");
    d.features_dc = reinterpret_cast<const int8_t *>(dc.data.data()); d.features_rest = reinterpret_cast<const int8_t *>(rest.data.data());
    d.scaling_scale = (float)scalar_or(m, "scaling_scale", 1.0); d.scaling_zero_point = (int32_t)scalar_or(m, "scaling_zero_point", 0);
    d.rotation_scale = (float)scalar_or(m, "rotation_scale", 1.0); d.rotation_zero_point = (int32_t)scalar_or(m, "rotation_zero_point", 0);
    auto quant = [&](ws_quantization &q, const std::string &key, bool present) {
        q.zero_point = present ? (int32_t)scalar_or(m, key + "_zero_point", 0) : 0;
        q.scale = present ? (float)scalar_or(m, key + "_scale", 1.0) : 1.f;
    };
    quant(d.quantization.color_dc, "features_dc", true); quant(d.quantization.color_rest, "features_rest", true);
    quant(d.quantization.opacity, "opacity", true); quant(d.quantization.scaling_factor, "scaling_factor", has_sf);
    if (const npz::Array *a = maybe(m, "mip_splatting")) { d.has_mip_splatting = 1; d.mip_splatting = a->as_double(0) != 0.0; }
    if (const npz::Array *a = maybe(m, "kernel_size")) { d.has_kernel_size = 1; d.kernel_size = (float)a->as_double(0); }
    if (const npz::Array *a = maybe(m, "background_color")) {
        if (a->count() < 3) throw std::runtime_error("background_color needs 3 entries");
        d.has_background = 1;
        for (int i = 0; i < 3; i++) d.background_color[i] = (float)a->as_double(i);
    }
    return d;
}

ws::PointCloud pointcloud_from_npz(const ws::Context &ctx, const std::string &file)
{
    const std::map<std::string, npz::Array> m = npz::read(reinterpret_cast<const uint8_t *>(file.data()), file.size());
    std::vector<int32_t> gi, fi;
    const ws_c3dgs_arrays d = c3dgs_from_members(m, gi, fi);
    return ws::PointCloud::from_c3dgs(ctx, d);
}

// ---- PNG (RGBA8, filter 0, stored deflate blocks) -----------------------------------------------------------------------
uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n)
{
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}
void put_be32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void put_chunk(std::vector<uint8_t> &out, const char tag[4], const std::vector<uint8_t> &data)
{
    put_be32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), tag, tag + 4); out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32_update(0, out.data() + start, out.size() - start));
}
std::vector<uint8_t> png_bytes(const uint8_t *rgba, uint32_t w, uint32_t h)
{
    std::vector<uint8_t> raw; raw.reserve((size_t)h * (1 + (size_t)w * 4));
    for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba + (size_t)y * w * 4, rgba + (size_t)(y + 1) * w * 4); }
    std::vector<uint8_t> z; z.push_back(0x78); z.push_back(0x01);        // zlib header, no compression
    uint32_t a = 1, b = 0;                                               // adler32 of `raw`
    for (size_t pos = 0; pos < raw.size() || pos == 0; ) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0); z.push_back(n & 0xff); z.push_back(n >> 8); z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; i++) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += n;
        if (last) break;
    }
    put_be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::vector<uint8_t> ihdr; put_be32(ihdr, w); put_be32(ihdr, h); ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});
    put_chunk(out, "IHDR", ihdr); put_chunk(out, "IDAT", z); put_chunk(out, "IEND", {});
    return out;
}
void write_png_from_f16(const std::string &path, const uint16_t *frame, uint32_t w, uint32_t h)
{
    std::vector<uint8_t> px((size_t)w * h * 4);
    for (size_t i = 0; i < px.size(); i++) px[i] = ws::half_to_u8(frame[i]);       // bin/render.rs:234-240
    const std::vector<uint8_t> png = png_bytes(px.data(), w, h);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f.write(reinterpret_cast<const char *>(png.data()), (std::streamsize)png.size());
}

// bin/render.rs:57-63: views wider than 1600 px are scaled down to 1600 (f32 arithmetic)
void render_resolution(uint32_t w, uint32_t h, uint32_t &ow, uint32_t &oh)
{
    ow = w; oh = h;
    if (w > 1600) { const float s = (float)w / 1600.f; ow = 1600; oh = (uint32_t)((float)h / s); }
}

// render_views (bin/render.rs:33-127)
void render_views(const ws::Context &ctx, ws::GaussianRenderer &renderer, const ws::PointCloud &pc, const std::vector<ws::SceneCamera> &cameras,
                  const std::string &img_out, const char *split)
{
    (void)ctx;
    const std::string dir = img_out + "/" + split;
    mkdir(img_out.c_str(), 0777); mkdir(dir.c_str(), 0777);
    std::printf("saving images to '%s'\n", dir.c_str());
    const ws::Aabb bbox = pc.bbox();
    std::vector<uint16_t> frame;
    for (size_t i = 0; i < cameras.size(); i++) {
        uint32_t w, h;
        render_resolution(cameras[i].width, cameras[i].height, w, h);
        ws::SplattingArgs args;
        args.camera = cameras[i].into_perspective();
        args.camera.fit_near_far(bbox);
        args.viewport = {w, h};
        args.gaussian_scaling = 1.f; args.max_sh_deg = pc.sh_deg(); args.walltime_secs = 100.f;     // bin/render.rs:92-104
        renderer.prepare(nullptr, pc, args);
        frame.assign((size_t)w * h * 4, 0);
        renderer.render_to_host(nullptr, pc, frame.data(), (size_t)w * 8, {0, 0, 0, 0});
        renderer.stats();                                                                            // synchronises the frame
        char name[32]; std::snprintf(name, sizeof name, "/%05zu.png", i);
        write_png_from_f16(dir + name, frame.data(), w, h);
    }
}

int usage()
{
    std::fprintf(stderr, "usage: ws_render <input.ply|.npz> <cameras.json> <img_out> [--max-sh-deg N]\n"
                         "       ws_render --parse-scene <cameras.json>\n"
                         "       ws_render --parse-npz <file.npz>\n"
                         "       ws_render --check-npz <file.npz>\n"
                         "       ws_render --png-selftest <out.png> <width> <height>\n");
    return 64;
}

}  // namespace

int main(int argc, char **argv)
{
    try {
        if (argc >= 3 && std::string(argv[1]) == "--parse-scene") {
            const ws::Scene scene = scene_from_json(read_file(argv[2]));
            std::printf("cameras %zu extend %.9g\n", scene.num_cameras(), scene.extend());
            for (const ws::SceneCamera &c : scene.cameras()) {
                const ws::PerspectiveCamera p = c.into_perspective();
                uint32_t w, h; render_resolution(c.width, c.height, w, h);
                std::printf("%zu %s %s %u %u %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", c.id, c.img_name.c_str(), ws::to_string(c.split), w, h,
                            p.position[0], p.position[1], p.position[2], p.rotation[0], p.rotation[1], p.rotation[2], p.rotation[3],
                            p.projection.fovx, p.projection.fovy, p.projection.fov2view_ratio);
            }
            return 0;
        }
        if (argc >= 3 && std::string(argv[1]) == "--parse-npz") {
            const std::string file = read_file(argv[2]);
            for (const auto &kv : npz::read(reinterpret_cast<const uint8_t *>(file.data()), file.size())) {
                uint64_t hsh = 1469598103934665603ull;                              // FNV-1a over the payload
                for (uint8_t b : kv.second.data) { hsh ^= b; hsh *= 1099511628211ull; }
                std::printf("%s %s", kv.first.c_str(), kv.second.descr.c_str());
                for (size_t s_ : kv.second.shape) std::printf(" %zu", s_);
                std::printf(" | %zu %016llx\n", kv.second.data.size(), (unsigned long long)hsh);
            }
            return 0;
        }
        if (argc >= 3 && std::string(argv[1]) == "--check-npz") {          // the ws_c3dgs_arrays the loader would hand to the GPU (no GPU)
            const std::string file = read_file(argv[2]);
            const std::map<std::string, npz::Array> m = npz::read(reinterpret_cast<const uint8_t *>(file.data()), file.size());
            std::vector<int32_t> gi, fi;
            const ws_c3dgs_arrays d = c3dgs_from_members(m, gi, fi);
            std::printf("points %llu covars %llu features %llu sh_deg %u scaling_factor %d gaussian_indices %d feature_indices %d\n",
                        (unsigned long long)d.num_points, (unsigned long long)d.num_covars, (unsigned long long)d.num_features, d.sh_deg,
                        d.scaling_factor != nullptr, d.gaussian_indices != nullptr, d.feature_indices != nullptr);
            std::printf("scaling %.9g %d rotation %.9g %d\n", d.scaling_scale, d.scaling_zero_point, d.rotation_scale, d.rotation_zero_point);
            const ws_quantization *q[4] = {&d.quantization.color_dc, &d.quantization.color_rest, &d.quantization.opacity, &d.quantization.scaling_factor};
            for (int i = 0; i < 4; i++) std::printf("quant %d %.9g\n", q[i]->zero_point, q[i]->scale);
            std::printf("meta mip %d %d kernel %d %.9g bg %d %.9g %.9g %.9g\n", d.has_mip_splatting, d.mip_splatting, d.has_kernel_size, d.kernel_size,
                        d.has_background, d.background_color[0], d.background_color[1], d.background_color[2]);
            if (d.gaussian_indices && d.num_points) std::printf("gi0 %d gilast %d\n", d.gaussian_indices[0], d.gaussian_indices[d.num_points - 1]);
            return 0;
        }
        if (argc >= 5 && std::string(argv[1]) == "--png-selftest") {
            const uint32_t w = (uint32_t)std::atoi(argv[3]), h = (uint32_t)std::atoi(argv[4]);
            std::vector<uint16_t> frame((size_t)w * h * 4);
            // a ramp of half bit patterns that covers negatives, subnormals, [0,1], > 1, inf and NaN
            for (size_t i = 0; i < frame.size(); i++) frame[i] = (uint16_t)((i * 2654435761u) >> 16);
            write_png_from_f16(argv[2], frame.data(), w, h);
            return 0;
        }
        if (argc < 4) return usage();
        uint32_t max_sh_deg = 3;
        for (int i = 4; i + 1 < argc; i++) if (std::string(argv[i]) == "--max-sh-deg") max_sh_deg = (uint32_t)std::atoi(argv[i + 1]);
        (void)max_sh_deg;                                    // parsed like the reference's Opt; render_views uses pc.sh_deg() (bin/render.rs:95)
        std::printf("reading scene file '%s'\n", argv[2]);
        const ws::Scene scene = scene_from_json(read_file(argv[2]));
        std::printf("reading point cloud file '%s'\n", argv[1]);
        const std::string file = read_file(argv[1]);
        // GenericGaussianPointCloud::load (io/mod.rs:44-61): dispatch on the magic bytes
        const bool is_ply = file.compare(0, 3, "ply") == 0, is_npz = file.compare(0, 4, "PK\x03\x04") == 0;
        if (!is_ply && !is_npz) throw std::runtime_error("Unknown file format");
        ws::Context ctx(0);
        ws::PointCloud pc = is_ply ? ws::PointCloud::from_ply(ctx, file.data(), file.size()) : pointcloud_from_npz(ctx, file);
        ws::GaussianRenderer renderer = ws::GaussianRenderer::new_(ctx, WS_FORMAT_RGBA16_FLOAT, pc.sh_deg(), pc.compressed());
        render_views(ctx, renderer, pc, scene.cameras(ws::Split::Test), argv[3], "test");
        render_views(ctx, renderer, pc, scene.cameras(ws::Split::Train), argv[3], "train");
        std::printf("done!\n");
        return 0;
    } catch (const ws::Error &e) {
        std::fprintf(stderr, "ws_render: %s\n", e.what());
        return 2;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ws_render: %s\n", e.what());
        return 1;
    }
}
