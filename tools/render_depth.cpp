// render_depth <cfg.json> — the reference tool (render/tools/render_depth.cpp:11-79) on the CUDA rasteriser.
// Same argument, same JSON keys and defaults (misc/render_depth.json), same call sequence on the same class
// surface (visma_b200::Renderer mirrors feh::Renderer) — including the tool's quirk of passing fy where cy belongs
// (:31) — and the same output files: <output_path>/depthmap.bin and mask.bin in feh::SaveMat's layout
// (core/utils.h:359-373: int32 rows, int32 cols, rows*cols values).  No OpenGL / GLFW / OpenCV / jsoncpp / libigl:
// the config reader below covers the JSON the tool's configs use (objects, arrays, numbers, strings, booleans,
// null, // and /* */ comments as jsoncpp accepts them) and the OBJ reader keeps the first three columns of `v`
// and `f` lines, as LoadMesh does (core/utils.cpp:125-135).  "show" needs a display and is ignored.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "renderer_b200.h"

namespace {

struct JsonValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JsonValue> arr;
    std::map<std::string, JsonValue> obj;

    // jsoncpp's cfg.get(key, default).asX()
    double number(const std::string &key, double def) const {
        auto it = obj.find(key);
        return it != obj.end() && it->second.kind == Number ? it->second.num : def;
    }
    bool boolean(const std::string &key, bool def) const {
        auto it = obj.find(key);
        return it != obj.end() && it->second.kind == Bool ? it->second.b : def;
    }
    std::string string(const std::string &key, const std::string &def) const {
        auto it = obj.find(key);
        return it != obj.end() && it->second.kind == String ? it->second.str : def;
    }
};

class JsonReader {
public:
    explicit JsonReader(const std::string &text) : s_(text) {}
    JsonValue parse() {
        JsonValue v = value();
        skip();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    [[noreturn]] void fail(const char *what) const {
        throw std::runtime_error("json: " + std::string(what) + " at offset " + std::to_string(i_));
    }
    void skip() {
        for (;;) {
            while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_;
            if (i_ + 1 < s_.size() && s_[i_] == '/' && s_[i_ + 1] == '/') {
                while (i_ < s_.size() && s_[i_] != '\n') ++i_;
            } else if (i_ + 1 < s_.size() && s_[i_] == '/' && s_[i_ + 1] == '*') {
                size_t e = s_.find("*/", i_ + 2);
                if (e == std::string::npos) fail("unterminated comment");
                i_ = e + 2;
            } else {
                return;
            }
        }
    }
    bool eat(char c) {
        skip();
        if (i_ < s_.size() && s_[i_] == c) { ++i_; return true; }
        return false;
    }
    std::string quoted() {
        if (!eat('"')) fail("expected string");
        std::string out;
        while (i_ < s_.size() && s_[i_] != '"') {
            char c = s_[i_++];
            if (c == '\\') {
                if (i_ >= s_.size()) fail("bad escape");
                char e = s_[i_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': fail("\\u escapes are not supported"); break;
                    default: out += e; break;  // \" \\ \/
                }
            } else {
                out += c;
            }
        }
        if (i_ >= s_.size()) fail("unterminated string");
        ++i_;
        return out;
    }
    JsonValue value() {
        skip();
        if (i_ >= s_.size()) fail("unexpected end");
        JsonValue v;
        const char c = s_[i_];
        if (c == '{') {
            ++i_;
            v.kind = JsonValue::Object;
            if (eat('}')) return v;
            do {
                skip();
                std::string k = quoted();
                if (!eat(':')) fail("expected ':'");
                v.obj[k] = value();
            } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            ++i_;
            v.kind = JsonValue::Array;
            if (eat(']')) return v;
            do v.arr.push_back(value()); while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            v.kind = JsonValue::String;
            v.str = quoted();
        } else if (s_.compare(i_, 4, "true") == 0) {
            v.kind = JsonValue::Bool; v.b = true; i_ += 4;
        } else if (s_.compare(i_, 5, "false") == 0) {
            v.kind = JsonValue::Bool; v.b = false; i_ += 5;
        } else if (s_.compare(i_, 4, "null") == 0) {
            i_ += 4;
        } else {
            const char *begin = s_.c_str() + i_;
            char *end = nullptr;
            v.num = std::strtod(begin, &end);
            if (end == begin) fail("unexpected character");
            v.kind = JsonValue::Number;
            i_ += (size_t)(end - begin);
        }
        return v;
    }
    const std::string &s_;
    size_t i_ = 0;
};

JsonValue LoadJson(const std::string &filename) {  // core/utils.cpp:148-157
    std::ifstream in(filename);
    if (!in.is_open()) throw std::runtime_error("failed to read file " + filename);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    return JsonReader(text).parse();
}

// `v x y z [...]` and `f a[/..] b[/..] c[/..] [...]` lines; 1-based (or negative = relative) indices
bool LoadObj(const std::string &file, std::vector<float> &V, std::vector<int> &F) {
    std::ifstream in(file);
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tag;
        ls >> tag;
        if (tag == "v") {
            float x, y, z;
            if (ls >> x >> y >> z) { V.push_back(x); V.push_back(y); V.push_back(z); }
        } else if (tag == "f") {
            std::string tok;
            int idx[3], k = 0;
            while (k < 3 && ls >> tok) {
                const int i = std::atoi(tok.c_str());  // stops at the first '/'
                idx[k++] = i > 0 ? i - 1 : (int)(V.size() / 3) + i;
            }
            if (k == 3) { F.push_back(idx[0]); F.push_back(idx[1]); F.push_back(idx[2]); }
        }
    }
    return !V.empty() && !F.empty();
}

template <typename T>
void SaveMat(const std::string &filename, const T *data, int rows, int cols) {  // core/utils.h:359-373
    std::ofstream out(filename, std::ios::out | std::ios::binary);
    if (!out.is_open()) throw std::runtime_error("failed to open " + filename);
    const int32_t hw[2] = {rows, cols};
    out.write(reinterpret_cast<const char *>(hw), sizeof hw);
    out.write(reinterpret_cast<const char *>(data), sizeof(T) * (size_t)rows * (size_t)cols);
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        std::cerr << "usage: render_depth YOUR_CONFIGURATION.json\n";
        return 2;
    }
    try {
        const JsonValue cfg = LoadJson(argv[1]);
        // the OpenGL context version: read for compatibility, meaningless here
        const int major_version = (int)cfg.number("major_version", 4);
        const int minor_version = (int)cfg.number("minor_version", 3);
        const int imH = (int)cfg.number("image_height", 480);
        const int imW = (int)cfg.number("image_width", 640);
        const float z_near = (float)cfg.number("z_near", 0.05);
        const float z_far = (float)cfg.number("z_far", 10.0);
        const float fx = (float)cfg.number("fx", 400);
        const float fy = (float)cfg.number("fy", 400);
        const float cx = (float)cfg.number("cx", 320);
        const float cy = (float)cfg.number("cy", 240);  // read and then unused, as in the tool
        (void)cy;
        visma_b200::Renderer::Mat4fc g_curr_init = visma_b200::Renderer::Mat4fc::Identity();

        auto ptr = std::make_shared<visma_b200::Renderer>(imH, imW, major_version, minor_version);
        ptr->SetCamera(z_near, z_far, fx, fy, cx, fy);  // render_depth.cpp:31 (sic)
        ptr->SetCamera(g_curr_init);
        std::vector<float> V;
        std::vector<int> F;
        const std::string mesh = cfg.string("mesh", "misc/hermanmiller_aeron.obj");
        if (!LoadObj(mesh, V, F)) throw std::runtime_error("failed to load mesh " + mesh);
        ptr->SetMesh(V, F);
        std::cout << "mesh " << mesh << ": " << V.size() / 3 << " vertices, " << F.size() / 3 << " faces\n";

        // SE3f(SO3f(), translation): identity rotation (:44)
        visma_b200::Renderer::Mat4fc model_pose = visma_b200::Renderer::Mat4fc::Identity();
        auto tr = cfg.obj.find("translation");
        if (tr == cfg.obj.end() || tr->second.kind != JsonValue::Array || tr->second.arr.size() != 3)
            throw std::runtime_error("\"translation\" must be an array of 3 numbers");
        for (int i = 0; i < 3; i++) model_pose(i, 3) = (float)tr->second.arr[i].num;

        std::vector<float> depth_map((size_t)imH * imW);
        ptr->RenderDepth(model_pose, depth_map.data());
        std::vector<uint8_t> mask;
        const bool has_mask = cfg.boolean("mask", false);
        if (has_mask) {
            mask.resize((size_t)imH * imW);
            ptr->RenderMask(model_pose, mask.data());
        }
        size_t covered = 0;
        for (float z : depth_map) covered += z < 1.0f;
        std::cout << "rendered " << imH << "x" << imW << ", " << covered << " covered pixels\n";

        if (cfg.boolean("save", false)) {
            const std::string output_path = cfg.string("output_path", ".");
            try {
                SaveMat<float>(output_path + "/depthmap.bin", depth_map.data(), imH, imW);
            } catch (const std::exception &) {
                std::cout << "failed to write out depth map to " << output_path << std::endl;
            }
            if (has_mask) {
                try {
                    SaveMat<uint8_t>(output_path + "/mask.bin", mask.data(), imH, imW);
                } catch (const std::exception &) {
                    std::cout << "failed to write out mask to " << output_path << std::endl;
                }
            }
        }
    } catch (const std::exception &e) {
        std::cerr << "render_depth: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
