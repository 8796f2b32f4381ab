// apps/triplet_match — working drop-in for the reference's stale CLI (apps/triplet_match.cpp:13-50):
//     triplet_match <model.pcd> <scene.pcd> [dist_thres=1.0] [model_match_factor=0.5] [max_icp_iterations=5] [curvature_test=1]
// Same positional contract (argv[1] model, argv[2] scene).  The reference app predates its own
// library (it calls scene::find with an external voxel_score functor); this one runs the HEAD
// API — model::init + scene::find_all_parallel — on the GPU and prints every accepted instance:
// inlier count, score and the model->scene pose.  Clouds are PCD files with the PointSurfel
// fields (tangent in radius/confidence/curvature, include/common:62-70).
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <triplet_match/scene>

namespace tr = triplet_match;
typedef pcl::PointSurfel point_t;
typedef tr::pointcloud<point_t> cloud_t;

int main(int argc, char const* argv[]) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <model.pcd> <scene.pcd> [dist_thres] [model_match_factor] [max_icp_iterations] [curvature_test]\n", argv[0]);
        return 2;
    }
    const float dist_thres = argc > 3 ? std::atof(argv[3]) : 1.0f;
    const float match_factor = argc > 4 ? std::atof(argv[4]) : 0.5f;
    const uint32_t icp_iters = argc > 5 ? static_cast<uint32_t>(std::atoi(argv[5])) : 5u;
    // the reference's pc_min/pc_max < 0.2 criterion (30-NN curvature); 0 for clouds with exact analytic normals
    const bool curvature_test = argc > 6 ? std::atoi(argv[6]) != 0 : true;
    try {
        cloud_t::Ptr model_cloud = cloud_t::from_pcd(argv[1]);
        tr::discretization_params dparams{20.f, 10.f / 180.f * static_cast<float>(M_PI), 10.f};
        std::cout << "init model (" << model_cloud->size() << " points)\n";
        tr::model<point_t> m(model_cloud, dparams);
        m.set_curvature_test(curvature_test);
        tr::sample_parameters sp{};
        sp.min_diameter_factor = 0.2f;
        sp.max_diameter_factor = 1.0f;
        sp.force_up = false;
        m.init(sp);
        std::cout << "init scene\n";
        cloud_t::Ptr scene_cloud = cloud_t::from_pcd(argv[2]);
        tr::scene<point_t> s(scene_cloud);
        s.set_curvature_test(curvature_test);
        std::cout << "start find (" << scene_cloud->size() << " scene points)\n";
        auto matches = s.find_all_parallel(m, dist_thres, match_factor, 0.9f, sp, icp_iters);
        std::cout << "accepted " << matches.size() << " transformations\n";
        for (const auto& mt : matches) {
            std::printf("match inliers %zu score %.9g T", mt.scene_corrs.size(), mt.signed_score);
            for (int i = 0; i < 16; ++i) std::printf(" %.9g", mt.transform.data()[i]);  // column-major, model -> scene
            std::printf("\n");
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "triplet_match: %s\n", e.what());
        return 1;
    }
    return 0;
}
