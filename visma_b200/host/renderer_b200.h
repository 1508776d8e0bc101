// renderer_b200.h — feh::Renderer's depth interface (render/renderer.h:41-118) without OpenGL / GLFW / GLEW /
// glm / glog / OpenCV: same constructor, SetCamera x3, SetMesh x3, RenderDepth, accessors; rendering runs in
// libvisma_b200.so's CUDA z-buffer rasteriser.  Header-only; needs only Eigen.
#pragma once

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>

#include "visma_b200.h"

namespace visma_b200 {

template <typename T>
T LinearizeDepth(T zb, T z_near, T z_far) {  // render/renderer.h:32-36
    return 2 * z_near * z_far / (z_far + z_near - (2 * zb - 1) * (z_far - z_near));
}

class Renderer {
public:
    typedef Eigen::Matrix<float, 4, 4, Eigen::ColMajor> Mat4fc;

    // major/minor were the OpenGL context version (render/renderer.cpp:48); accepted for source compatibility
    Renderer(int maxHeight, int maxWidth, int major_version = 3, int minor_version = 2, int device = 0)
        : rows_(maxHeight), cols_(maxWidth), device_(device) {
        (void)major_version; (void)minor_version;
        pose_.setIdentity();
    }

    void SetCamera(float z_near, float z_far, const float *intrinsics) {
        z_near_ = z_near; z_far_ = z_far;
        fx_ = intrinsics[0]; fy_ = intrinsics[1]; cx_ = intrinsics[2]; cy_ = intrinsics[3];
        // the reference uploads view = vision_to_graphics here (render/renderer.cpp:270-276), which discards
        // any pose set earlier: intrinsics first, pose second
        pose_.setIdentity();
    }
    void SetCamera(float z_near, float z_far, float fx, float fy, float cx, float cy) {
        float k[4] = {fx, fy, cx, cy};
        SetCamera(z_near, z_far, k);
    }
    /// pose: transformation from the initial camera frame to the current camera frame
    void SetCamera(const Mat4fc &pose) { pose_ = pose; }

    void SetMesh(float *vertices, int num_vertices, int *faces, int num_faces) {
        V_.assign(vertices, vertices + 3 * (size_t)num_vertices);
        F_.assign(faces, faces + 3 * (size_t)num_faces);
    }
    void SetMesh(const std::vector<float> &vertices, const std::vector<int> &faces) {
        V_ = vertices; F_ = faces;
    }
    void SetMesh(Eigen::Matrix<float, Eigen::Dynamic, 3, Eigen::RowMajor> vertices,
                 Eigen::Matrix<int, Eigen::Dynamic, 3, Eigen::RowMajor> faces) {
        SetMesh(vertices.data(), (int)vertices.rows(), faces.data(), (int)faces.rows());
    }

    /// out: rows x cols floats, window depth in [0,1], 1 = background, row 0 = image top
    /// (what glReadPixels(GL_DEPTH_COMPONENT, GL_FLOAT) returned, render/renderer.cpp:343)
    void RenderDepth(const Mat4fc &model, float *out) { RenderDepthBatch(&model, 1, out); }

    /// n poses of the current mesh in one launch sequence; out: n x rows x cols floats
    void RenderDepthBatch(const Mat4fc *models, int n, float *out) {
        std::vector<float> V((size_t)n * V_.size()), M(16 * (size_t)n);
        std::vector<int32_t> F((size_t)n * F_.size());
        std::vector<int64_t> voff(n + 1), foff(n + 1);
        for (int i = 0; i < n; i++) {
            std::copy(V_.begin(), V_.end(), V.begin() + (size_t)i * V_.size());
            std::copy(F_.begin(), F_.end(), F.begin() + (size_t)i * F_.size());
            voff[i] = (int64_t)i * (V_.size() / 3);
            foff[i] = (int64_t)i * (F_.size() / 3);
            std::memcpy(M.data() + 16 * i, models[i].data(), 16 * sizeof(float));
        }
        voff[n] = (int64_t)n * (V_.size() / 3);
        foff[n] = (int64_t)n * (F_.size() / 3);
        int rc = vb200_render_depth_batch(V.data(), voff.data(), F.data(), foff.data(), n, M.data(), pose_.data(),
                                          z_near_, z_far_, fx_, fy_, cx_, cy_, rows_, cols_, device_, nullptr, out);
        if (rc != VB200_OK)  // the reference aborts through glog CHECK / LOG(FATAL) on renderer failures
            throw std::runtime_error(std::string("visma_b200::Renderer: ") + vb200_strerror(rc) + " " +
                                     vb200_last_error());
    }

    /// RenderEdge / RenderMask (render/renderer.h:82-97): out = rows x cols bytes.  The edge pass linearises
    /// depth with the shader's own fixed uniforms z_near = 0.05, z_far = 2.0 (render/renderer.cpp:95-96).  The
    /// mask keeps the reference's polarity: 255 on background (its clear colour, :411-422), 0 on the mesh.
    void RenderEdge(const Mat4fc &model, uint8_t *out) { EdgeMask(model, out, nullptr); }
    void RenderMask(const Mat4fc &model, uint8_t *out) { EdgeMask(model, nullptr, out); }

    float fx() const { return fx_; }
    float fy() const { return fy_; }
    float cx() const { return cx_; }
    float cy() const { return cy_; }
    float z_near() const { return z_near_; }
    float z_far() const { return z_far_; }
    int width() const { return cols_; }
    int cols() const { return cols_; }
    int height() const { return rows_; }
    int rows() const { return rows_; }

private:
    void EdgeMask(const Mat4fc &model, uint8_t *edge, uint8_t *mask) {
        int64_t voff[2] = {0, (int64_t)(V_.size() / 3)}, foff[2] = {0, (int64_t)(F_.size() / 3)};
        int rc = vb200_render_edge_mask_batch(V_.data(), voff, F_.data(), foff, 1, model.data(), pose_.data(), z_near_,
                                              z_far_, fx_, fy_, cx_, cy_, rows_, cols_, device_, 0.05f, 2.0f, edge,
                                              mask, 0);
        if (rc != VB200_OK)
            throw std::runtime_error(std::string("visma_b200::Renderer: ") + vb200_strerror(rc) + " " +
                                     vb200_last_error());
    }

    int rows_, cols_, device_;
    float fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0, z_near_ = 0.05f, z_far_ = 10.0f;
    Mat4fc pose_;
    std::vector<float> V_;
    std::vector<int> F_;
};

}  // namespace visma_b200
