// Drives EVERY visible GPU through the C-ABI from one process: pigo_init_devices(mask) +
// pigo_run_cascade_batch_sharded / pigo_detect_batch_sharded via the C++ mirror, and checks that the sharded results are
// byte-identical to the single-device calls (frame order restored, replicas built on first use, generator keys independent
// of the split).  Exit code 0 = ok; prints "devices=N" so the caller can require N >= 2.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "pigo_b200.hpp"

static std::vector<uint8_t> slurp(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  const std::string root = argc > 1 ? argv[1] : ".";
  const auto gray = slurp(root + "/tests/golden/sample_gray_400x320.u8");
  if (gray.size() != 400 * 320) { std::fprintf(stderr, "fixtures missing\n"); return 2; }
  try {
    const int ndev = pigo_device_count();
    std::printf("devices=%d\n", ndev);
    if (ndev < 1) return 2;
    // 9 frames of 540x960: the sample face tiled with a per-frame shift
    const int R = 540, C = 960, NF = 9;
    std::vector<uint8_t> frames((size_t)NF * R * C);
    for (int f = 0; f < NF; ++f)
      for (int y = 0; y < R; ++y)
        for (int x = 0; x < C; ++x) frames[((size_t)f * R + y) * C + x] = gray[(size_t)((y + 37 * f) % 400) * 320 + (x + 53 * f) % 320];
    pigo::Pigo face = pigo::Pigo::Unpack(slurp(root + "/pigo_b200/data/cascade/facefinder"));
    pigo::PuplocCascade plc = pigo::PuplocCascade::UnpackCascade(slurp(root + "/pigo_b200/data/cascade/puploc"));
    const char* names[9] = {"lp46", "lp44", "lp42", "lp38", "lp312", "lp93", "lp84", "lp82", "lp81"};
    std::vector<pigo::PuplocCascade> flp;
    for (auto n : names) flp.push_back(pigo::PuplocCascade::UnpackFlp(root + "/pigo_b200/data/cascade/lps/" + n));
    std::vector<std::pair<const pigo::PuplocCascade*, bool>> calls;   // core/flploc_test.go:122-146
    for (int e = 0; e < 5; ++e) { calls.push_back({&flp[e], false}); calls.push_back({&flp[e], true}); }
    for (int m = 5; m < 9; ++m) calls.push_back({&flp[m], false});
    calls.push_back({&flp[6], true});
    pigo::CascadeParams cp;
    cp.Image = pigo::ImageParams{nullptr, R, C, C};
    cp.MinSize = 20; cp.MaxSize = 1000; cp.ShiftFactor = 0.2; cp.ScaleFactor = 1.1;

    pigo::InitDevices(1u);                                   // single device: the reference result
    auto one = face.RunCascadeBatch(frames.data(), NF, (size_t)R * C, cp, 0.0, false);
    auto pone = pigo::DetectBatch(face, plc, calls, frames.data(), NF, (size_t)R * C, cp, 0.1, 50, 50, 63, 11, false);
    pigo::InitDevices((1u << ndev) - 1u);                    // all devices
    auto all = face.RunCascadeBatch(frames.data(), NF, (size_t)R * C, cp, 0.0, true);
    auto pall = pigo::DetectBatch(face, plc, calls, frames.data(), NF, (size_t)R * C, cp, 0.1, 50, 50, 63, 11, true);
    size_t ndet = 0, nref = 0;
    for (int f = 0; f < NF; ++f) {
      if (one[f].size() != all[f].size()) { std::printf("frame %d: %zu vs %zu detections\n", f, one[f].size(), all[f].size()); return 1; }
      for (size_t i = 0; i < one[f].size(); ++i)
        if (std::memcmp(&one[f][i], &all[f][i], sizeof(pigo::Detection)) != 0) { std::printf("frame %d det %zu differs\n", f, i); return 1; }
      ndet += one[f].size();
      if (pone[f].size() != pall[f].size()) { std::printf("frame %d: face count differs\n", f); return 1; }
      for (size_t k = 0; k < pone[f].size(); ++k) {
        const auto &a = pone[f][k], &b = pall[f][k];
        if (std::memcmp(&a.Face, &b.Face, sizeof(pigo::Detection)) != 0 || a.Refined != b.Refined) return 1;
        if (!a.Refined) continue;
        ++nref;
        if (a.LeftEye.Row != b.LeftEye.Row || a.LeftEye.Col != b.LeftEye.Col || a.RightEye.Row != b.RightEye.Row || a.Landmarks.size() != 15) return 1;
        for (size_t c = 0; c < 15; ++c)
          if (a.Landmarks[c].Row != b.Landmarks[c].Row || a.Landmarks[c].Col != b.Landmarks[c].Col || a.Landmarks[c].Scale != b.Landmarks[c].Scale) return 1;
      }
    }
    std::printf("sharded over %d device(s) == single device: %zu detections, %zu refined faces x 17 points\n", ndev, ndet, nref);
    return (ndet > 0 && nref > 0) ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
