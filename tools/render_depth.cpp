// render_depth <cfg.json> — the reference tool (render/tools/render_depth.cpp:11-79) on the CUDA rasteriser.
// Same argument, same JSON keys and defaults (misc/render_depth.json, read with the reference's own jsoncpp), same
// call sequence on the same class surface (visma_b200::Renderer mirrors feh::Renderer) — including the tool's quirk
// of passing fy where cy belongs (:31) — and the same output files: <output_path>/depthmap.bin and mask.bin in
// feh::SaveMat's layout (core/utils.h:359-373: int32 rows, int32 cols, rows*cols values).  No OpenGL / GLFW / OpenCV
// / libigl.  "show" needs a display and is ignored.
#include <memory>

#include "renderer_b200.h"
#include "tool_io.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::cerr << "usage: render_depth YOUR_CONFIGURATION.json\n";
        return 2;
    }
    try {
        const Json::Value cfg = tool_io::LoadJson(argv[1]);
        // the OpenGL context version: read for compatibility, meaningless here
        const int major_version = cfg.get("major_version", 4).asInt();
        const int minor_version = cfg.get("minor_version", 3).asInt();
        const int imH = cfg.get("image_height", 480).asInt();
        const int imW = cfg.get("image_width", 640).asInt();
        const float z_near = cfg.get("z_near", 0.05).asFloat();
        const float z_far = cfg.get("z_far", 10.0).asFloat();
        const float fx = cfg.get("fx", 400).asFloat();
        const float fy = cfg.get("fy", 400).asFloat();
        const float cx = cfg.get("cx", 320).asFloat();
        const float cy = cfg.get("cy", 240).asFloat();  // read and then unused, as in the tool
        (void)cy;
        visma_b200::Renderer::Mat4fc g_curr_init = visma_b200::Renderer::Mat4fc::Identity();

        auto ptr = std::make_shared<visma_b200::Renderer>(imH, imW, major_version, minor_version);
        ptr->SetCamera(z_near, z_far, fx, fy, cx, fy);  // render_depth.cpp:31 (sic)
        ptr->SetCamera(g_curr_init);
        std::vector<float> V;
        std::vector<int> F;
        const std::string mesh = cfg.get("mesh", "misc/hermanmiller_aeron.obj").asString();
        if (!tool_io::LoadObj(mesh, V, F)) throw std::runtime_error("failed to load mesh " + mesh);
        ptr->SetMesh(V, F);
        std::cout << "mesh " << mesh << ": " << V.size() / 3 << " vertices, " << F.size() / 3 << " faces\n";

        // SE3f(SO3f(), translation): identity rotation (:44)
        visma_b200::Renderer::Mat4fc model_pose = visma_b200::Renderer::Mat4fc::Identity();
        const Eigen::Matrix<float, 3, 1> tr = tool_io::GetMatrixFromJson<float, 3, 1>(cfg, "translation");
        for (int i = 0; i < 3; i++) model_pose(i, 3) = tr(i);

        std::vector<float> depth_map((size_t)imH * imW);
        ptr->RenderDepth(model_pose, depth_map.data());
        std::vector<uint8_t> mask;
        const bool has_mask = cfg.get("mask", false).asBool();
        if (has_mask) {
            mask.resize((size_t)imH * imW);
            ptr->RenderMask(model_pose, mask.data());
        }
        size_t covered = 0;
        for (float z : depth_map) covered += z < 1.0f;
        std::cout << "rendered " << imH << "x" << imW << ", " << covered << " covered pixels\n";

        if (cfg.get("save", false).asBool()) {
            const std::string output_path = cfg.get("output_path", ".").asString();
            try {
                tool_io::SaveMat<float>(output_path + "/depthmap.bin", depth_map.data(), imH, imW);
            } catch (const std::exception &) {
                std::cout << "failed to write out depth map to " << output_path << std::endl;
            }
            if (has_mask) {
                try {
                    tool_io::SaveMat<uint8_t>(output_path + "/mask.bin", mask.data(), imH, imW);
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
