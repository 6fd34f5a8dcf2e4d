// Replays the reference's own test (core/pigo_test.go:68-84: sample image -> RunCascade -> ClusterDetections(0.1) -> at
// least one face) through the C++ mirror of the Go API, and compares RunCascade with the committed oracle vector.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "pigo_b200.hpp"

static std::vector<uint8_t> slurp(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  const std::string root = argc > 1 ? argv[1] : ".";
  const auto casc = slurp(root + "/pigo_b200/data/cascade/facefinder");
  const auto gray = slurp(root + "/tests/golden/sample_gray_400x320.u8");
  if (casc.empty() || gray.size() != 400 * 320) { std::fprintf(stderr, "fixtures missing\n"); return 2; }
  try {
    pigo::Pigo p = pigo::Pigo::Unpack(casc);
    pigo::CascadeParams cp;
    cp.Image = pigo::ImageParams{gray.data(), 400, 320, 320};
    cp.MinSize = 20; cp.MaxSize = 1000; cp.ShiftFactor = 0.2; cp.ScaleFactor = 1.1;   // core/pigo_test.go:44-50
    auto dets = p.RunCascade(cp, 0.0);
    auto cl = p.ClusterDetections(dets, 0.1);
    std::printf("dets=%zu clusters=%zu", dets.size(), cl.size());
    for (auto& d : cl) std::printf(" (%d,%d,%d,%.4f)", d.Row, d.Col, d.Scale, d.Q);
    std::printf("\n");
    // oracle-generated vector for this buffer (tests/golden/oracle_vectors.npz: sample_test_dets)
    const int exp[4][3] = {{194, 151, 215}, {213, 166, 236}, {199, 143, 284}, {219, 157, 312}};
    if (dets.size() != 4 || cl.empty()) return 1;
    for (int i = 0; i < 4; ++i)   // dets are now sorted by Q ascending (ClusterDetections sorts in place)
      if (dets[i].Row != exp[i][0] || dets[i].Col != exp[i][1] || dets[i].Scale != exp[i][2]) return 1;
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
