// Generates the seeded input fixtures the reference's own tests use, with the same generator the
// reference uses (default-seeded std::mt19937_64 + std::uniform_real_distribution<float>,
// libstdc++), plus the host-side expectations those tests compute.
//   tests/test_cases/runtime/messaging/test_spatial_3d.cu:72-205 (Mandatory, 2049 agents in [0,5)^3)
//   tests/test_cases/runtime/messaging/test_spatial_2d.cu  (Mandatory twin, [0,11)^2 ... see below)
//   examples/cpp/circles_spatial3D/src/main.cu:174-182 (Circles initial population)
// Build + run:  g++ -O2 -o /tmp/gen_golden tests/golden/gen_golden.cpp && /tmp/gen_golden tests/golden
#include <cstdint>
#include <cstdio>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

static void write_f32(const std::string &path, const std::vector<float> &v) {
  FILE *f = fopen(path.c_str(), "wb");
  fwrite(v.data(), sizeof(float), v.size(), f);
  fclose(f);
}
static void write_u32(const std::string &path, const std::vector<uint32_t> &v) {
  FILE *f = fopen(path.c_str(), "wb");
  fwrite(v.data(), sizeof(uint32_t), v.size(), f);
  fclose(f);
}

int main(int argc, char **argv) {
  const std::string dir = argc > 1 ? argv[1] : ".";
  {  // Spatial3DMessageTest.Mandatory: positions, myBin, expected Moore-neighbourhood count per agent
    const int N = 2049;
    std::mt19937_64 rng;
    std::uniform_real_distribution<float> dist(0.0f, 5.0f);
    std::vector<float> pos(3 * N);
    std::vector<uint32_t> my_bin(N), expect(N);
    std::unordered_map<int, unsigned int> bin_counts;
    for (int i = 0; i < N; ++i) {
      float p[3] = {dist(rng), dist(rng), dist(rng)};
      pos[i] = p[0]; pos[N + i] = p[1]; pos[2 * N + i] = p[2];
      const unsigned int bp[3] = {(unsigned int)(p[0] / 1), (unsigned int)(p[1] / 1), (unsigned int)(p[2] / 1)};
      my_bin[i] = bp[2] * 25 + bp[1] * 5 + bp[0];
      bin_counts[my_bin[i]] += 1;
    }
    std::unordered_map<int, unsigned int> res;
    for (int x1 = 0; x1 < 5; ++x1) for (int y1 = 0; y1 < 5; ++y1) for (int z1 = 0; z1 < 5; ++z1) {
      unsigned int sum = 0;
      for (int x2 = -1; x2 <= 1; ++x2) for (int y2 = -1; y2 <= 1; ++y2) for (int z2 = -1; z2 <= 1; ++z2) {
        int b[3] = {x1 + x2, y1 + y2, z1 + z2};
        if (b[0] >= 0 && b[1] >= 0 && b[2] >= 0 && b[0] < 5 && b[1] < 5 && b[2] < 5) sum += bin_counts[b[2] * 25 + b[1] * 5 + b[0]];
      }
      res[z1 * 25 + y1 * 5 + x1] = sum;
    }
    for (int i = 0; i < N; ++i) expect[i] = res[my_bin[i]];
    write_f32(dir + "/mandatory3d_pos.f32", pos);
    write_u32(dir + "/mandatory3d_mybin.u32", my_bin);
    write_u32(dir + "/mandatory3d_expect.u32", expect);
  }
  {  // Spatial2DMessageTest.Mandatory (test_spatial_2d.cu:66-180): 2049 agents in [0,11)^2, radius 1, 11x11 bins
    const int N = 2049;
    std::mt19937_64 rng;
    std::uniform_real_distribution<float> dist(0.0f, 11.0f);
    std::vector<float> pos(2 * N);
    std::vector<uint32_t> my_bin(N), expect(N);
    std::unordered_map<int, unsigned int> bin_counts;
    for (int i = 0; i < N; ++i) {
      float p[2] = {dist(rng), dist(rng)};
      pos[i] = p[0]; pos[N + i] = p[1];
      const unsigned int bp[2] = {(unsigned int)(p[0] / 1), (unsigned int)(p[1] / 1)};
      my_bin[i] = bp[1] * 11 + bp[0];
      bin_counts[my_bin[i]] += 1;
    }
    std::unordered_map<int, unsigned int> res;
    for (int x1 = 0; x1 < 11; ++x1) for (int y1 = 0; y1 < 11; ++y1) {
      unsigned int sum = 0;
      for (int x2 = -1; x2 <= 1; ++x2) for (int y2 = -1; y2 <= 1; ++y2) {
        int b[2] = {x1 + x2, y1 + y2};
        if (b[0] >= 0 && b[1] >= 0 && b[0] < 11 && b[1] < 11) sum += bin_counts[b[1] * 11 + b[0]];
      }
      res[y1 * 11 + x1] = sum;
    }
    for (int i = 0; i < N; ++i) expect[i] = res[my_bin[i]];
    write_f32(dir + "/mandatory2d_pos.f32", pos);
    write_u32(dir + "/mandatory2d_mybin.u32", my_bin);
    write_u32(dir + "/mandatory2d_expect.u32", expect);
  }
  {  // Circles example initial population (N=16384, ENV_MAX=floor(cbrt(N))=25): x,y,z drawn per agent
    const int N = 16384;
    std::mt19937_64 rng;
    std::uniform_real_distribution<float> dist(0.0f, 25.0f);
    std::vector<float> pos(3 * N);
    for (int i = 0; i < N; ++i) { pos[i] = dist(rng); pos[N + i] = dist(rng); pos[2 * N + i] = dist(rng); }
    write_f32(dir + "/circles16k_pos.f32", pos);
  }
  return 0;
}
