// tool_io.h — the file IO the two harness tools share: JSON through the reference's own vendored jsoncpp
// (thirdparty/jsoncpp: comments in cfg/tool.json and misc/render_depth.json are accepted by its reader, as in
// core/utils.cpp:148-157), OBJ meshes (the first three columns of `v` and `f` lines, as LoadMesh / igl::readOBJ
// deliver them, core/utils.cpp:125-135), PLY point clouds (what Open3D's ReadPointCloudFromPLY takes from
// test.klg.ply: x y z [nx ny nz], ascii or binary_little_endian, any scalar property types), matrices in JSON as
// flat row-major lists (core/utils.h:300-339) and feh::SaveMat's binary layout (core/utils.h:359-373).
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <json/json.h>

namespace tool_io {

inline Json::Value LoadJson(const std::string &filename) {  // core/utils.cpp:148-157
    std::ifstream in(filename, std::ios::in);
    if (!in.is_open()) throw std::runtime_error("failed to read file " + filename);
    Json::Value out;
    in >> out;
    return out;
}

inline void SaveJson(const Json::Value &value, const std::string &filename) {  // core/utils.cpp:159-166
    std::ofstream out(filename, std::ios::out);
    if (!out.is_open()) throw std::runtime_error("failed to write file " + filename);
    out << value;
}

template <typename T, int R, int C>
Eigen::Matrix<T, R, C> GetMatrixFromJson(const Json::Value &v, const std::string &key) {  // core/utils.h:300-318
    Eigen::Matrix<T, R, C> m;
    if (!v.isMember(key) || (int)v[key].size() != R * C) throw std::runtime_error("bad matrix \"" + key + "\" in json");
    for (int i = 0; i < R; i++)
        for (int j = 0; j < C; j++) m(i, j) = (T)v[key][i * C + j].asDouble();
    return m;
}

template <typename Derived>
void WriteMatrixToJson(Json::Value &d, const std::string &key, const Eigen::MatrixBase<Derived> &m) {  // :333-339
    d[key] = Json::Value(Json::arrayValue);
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j) d[key].append(Json::Value((double)m(i, j)));
}

// `v x y z [...]` and `f a[/..] b[/..] c[/..] [...]` lines; 1-based (or negative = relative) indices
template <typename Real>
bool LoadObj(const std::string &file, std::vector<Real> &V, std::vector<int> &F) {
    std::ifstream in(file);
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tag;
        ls >> tag;
        if (tag == "v") {
            Real x, y, z;
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

// PLY vertices -> points (and normals when the file has nx ny nz)
inline bool LoadPly(const std::string &file, std::vector<Eigen::Vector3d> &points, std::vector<Eigen::Vector3d> &normals) {
    std::ifstream in(file, std::ios::binary);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    if (line.substr(0, 3) != "ply") return false;
    struct Prop { std::string name; int size; char kind; };  // kind: f float, d double, i signed, u unsigned
    std::vector<Prop> props;
    std::string fmt;
    size_t n_vertex = 0;
    bool in_vertex = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line == "end_header") break;
        std::istringstream ls(line);
        std::string a, b, c;
        ls >> a >> b >> c;
        if (a == "format") fmt = b;
        else if (a == "element") { in_vertex = b == "vertex"; if (in_vertex) n_vertex = (size_t)std::atoll(c.c_str()); }
        else if (a == "property" && in_vertex) {
            if (b == "list") return false;
            Prop p;
            p.name = c;
            if (b == "float" || b == "float32") { p.size = 4; p.kind = 'f'; }
            else if (b == "double" || b == "float64") { p.size = 8; p.kind = 'd'; }
            else if (b == "char" || b == "int8") { p.size = 1; p.kind = 'i'; }
            else if (b == "uchar" || b == "uint8") { p.size = 1; p.kind = 'u'; }
            else if (b == "short" || b == "int16") { p.size = 2; p.kind = 'i'; }
            else if (b == "ushort" || b == "uint16") { p.size = 2; p.kind = 'u'; }
            else if (b == "int" || b == "int32") { p.size = 4; p.kind = 'i'; }
            else if (b == "uint" || b == "uint32") { p.size = 4; p.kind = 'u'; }
            else return false;
            props.push_back(p);
        }
    }
    int col[6] = {-1, -1, -1, -1, -1, -1};
    const char *names[6] = {"x", "y", "z", "nx", "ny", "nz"};
    for (size_t i = 0; i < props.size(); i++)
        for (int k = 0; k < 6; k++)
            if (props[i].name == names[k]) col[k] = (int)i;
    if (col[0] < 0 || col[1] < 0 || col[2] < 0) return false;
    const bool has_n = col[3] >= 0 && col[4] >= 0 && col[5] >= 0;
    points.resize(n_vertex);
    normals.resize(has_n ? n_vertex : 0);
    std::vector<double> row(props.size());
    if (fmt == "ascii") {
        for (size_t v = 0; v < n_vertex; v++) {
            for (size_t i = 0; i < props.size(); i++) in >> row[i];
            for (int a = 0; a < 3; a++) points[v][a] = row[col[a]];
            if (has_n) for (int a = 0; a < 3; a++) normals[v][a] = row[col[3 + a]];
        }
    } else if (fmt == "binary_little_endian") {
        size_t stride = 0;
        for (auto &p : props) stride += p.size;
        std::vector<char> buf(stride * n_vertex);
        in.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)in.gcount() != buf.size()) return false;
        for (size_t v = 0; v < n_vertex; v++) {
            const char *p = buf.data() + v * stride;
            for (size_t i = 0; i < props.size(); i++) {
                const Prop &pr = props[i];
                double x = 0;
                if (pr.kind == 'f') { float t; std::memcpy(&t, p, 4); x = t; }
                else if (pr.kind == 'd') { std::memcpy(&x, p, 8); }
                else if (pr.size == 1) x = pr.kind == 'i' ? (double)*(const int8_t *)p : (double)*(const uint8_t *)p;
                else if (pr.size == 2) { int16_t t; std::memcpy(&t, p, 2); x = pr.kind == 'i' ? (double)t : (double)(uint16_t)t; }
                else { int32_t t; std::memcpy(&t, p, 4); x = pr.kind == 'i' ? (double)t : (double)(uint32_t)t; }
                row[i] = x;
                p += pr.size;
            }
            for (int a = 0; a < 3; a++) points[v][a] = row[col[a]];
            if (has_n) for (int a = 0; a < 3; a++) normals[v][a] = row[col[3 + a]];
        }
    } else {
        return false;
    }
    return true;
}

template <typename T>
void SaveMat(const std::string &filename, const T *data, int rows, int cols) {  // core/utils.h:359-373
    std::ofstream out(filename, std::ios::out | std::ios::binary);
    if (!out.is_open()) throw std::runtime_error("failed to open " + filename);
    const int32_t hw[2] = {rows, cols};
    out.write(reinterpret_cast<const char *>(hw), sizeof hw);
    out.write(reinterpret_cast<const char *>(data), sizeof(T) * (size_t)rows * (size_t)cols);
}

}  // namespace tool_io
