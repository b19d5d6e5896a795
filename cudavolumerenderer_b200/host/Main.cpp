// Main.cpp -- cvr_render: the reference's command line over the B200 path.
//
// Flag surface of ConfigParser.cpp:13-41 (scene-file/-s positional, --scene-type,
// --interactive, --trials, --algorithm/-a, --kernel/-k, --number-of-tiles X [Y],
// --use-unified-memory, --iterations/-i, --output/-o, --resolution/-r W [H]) without
// boost, Config::createConfig's scene-type auto-detection (:84-108) and the bench loop of
// Main.cpp:46-121 (first trial discarded, mean / std, "paths per sec").  Additions:
// --device N, --option key=value (forwarded to cvr_set_option), "synth:<name>" scenes.
// The interactive viewer (InteractiveRenderer.h) is out of scope (no display/GL).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "CudaVolPath.h"
#include "SceneBuilders.h"

using namespace cvrhost;

namespace {
const char* print_prefix = "[config] ";

struct Options {
  std::string scene_file, scene_type = "Auto", algorithm = "cudaVolPath", kernel = "regenerationSK", output;
  bool interactive = true;  // ConfigParser.cpp:18 default (Q19)
  unsigned trials = 1, iterations = 20;
  std::vector<unsigned> n_tiles{1, 1}, resolution{1024, 1024};
  bool unified_memory = false;
  int device = 0;
  int gpus = 1;                  // --gpus N: devices 0..N-1 of this process behind one device group
  std::string shard = "balanced"; // --shard tiles | spp | balanced
  bool use_group = false;        // --shard given: go through the device group also with one GPU
  bool dump_scene = false;  // print the loaded scene and exit (no GPU needed)
  std::string dump_raw;     // also write the float4 image as raw little-endian floats
  bool png = false;         // also write <output>.png (Image::savePNG, Image.cpp:35-56)
  std::vector<std::pair<std::string, std::string>> lib_options;
};

void usage() {
  std::cout << "Generic:\n"
               "  -h [ --help ]                 produce help message\n"
               "  -s [ --scene-file ] arg       scene file to parse (or synth:bucky|hetvol|manix|fbm[:n])\n"
               "  --scene-type arg (=Auto)      Auto | MitsubaXml | Vdb | Raw\n"
               "  --interactive arg (=1)        run the interactive view (not available in this build)\n"
               "  --trials arg (=1)             number of times to run the algorithm\n"
               "  -a [ --algorithm ] arg (=cudaVolPath)\n"
               "  -k [ --kernel ] arg (=regenerationSK)   naiveSK | regenerationSK | streamingSK | streamingMK | sortingSK\n"
               "  --number-of-tiles arg (=1 1)\n"
               "  --use-unified-memory arg (=0)\n"
               "  --device arg (=0)             CUDA device\n"
               "  --gpus arg (=1)               render on devices 0..N-1 (volume replicated, one NCCL reduce of the framebuffer)\n"
               "  --shard arg (=balanced)       tiles | spp | balanced: how the work is split over the GPUs\n"
               "  --option key=value            forwarded to cvr_set_option (rng, sched, layout, ...)\n"
               "Scene configuration override:\n"
               "  -i [ --iterations ] arg (=20)\n"
               "  -o [ --output ] arg\n"
               "  -r [ --resolution ] arg (=1024 1024)\n";
}

bool is_number(const char* s) { return s && *s && std::all_of(s, s + strlen(s), [](char c) { return c >= '0' && c <= '9'; }); }

bool parse(int argc, char** argv, Options& o) {
  auto need = [&](int& i) -> const char* {
    if (i + 1 >= argc) throw std::runtime_error(std::string("the required argument for option '") + argv[i] + "' is missing");
    return argv[++i];
  };
  auto as_bool = [](const std::string& v) { return v == "1" || v == "true" || v == "on" || v == "yes"; };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-h" || a == "--help") {
      usage();
      return false;
    } else if (a == "-s" || a == "--scene-file")
      o.scene_file = need(i);
    else if (a == "--scene-type")
      o.scene_type = need(i);
    else if (a == "--interactive")
      o.interactive = as_bool(need(i));
    else if (a == "--trials")
      o.trials = (unsigned)std::stoul(need(i));
    else if (a == "-a" || a == "--algorithm")
      o.algorithm = need(i);
    else if (a == "-k" || a == "--kernel")
      o.kernel = need(i);
    else if (a == "--use-unified-memory")
      o.unified_memory = as_bool(need(i));
    else if (a == "--device")
      o.device = std::stoi(need(i));
    else if (a == "--gpus")
      o.gpus = std::stoi(need(i));
    else if (a == "--shard")
      o.shard = need(i), o.use_group = true;
    else if (a == "-i" || a == "--iterations")
      o.iterations = (unsigned)std::stoul(need(i));
    else if (a == "-o" || a == "--output")
      o.output = need(i);
    else if (a == "--dump-scene")
      o.dump_scene = true;
    else if (a == "--dump-raw")
      o.dump_raw = need(i);
    else if (a == "--png")
      o.png = true;
    else if (a == "--option") {
      std::string kv = need(i);
      size_t eq = kv.find('=');
      if (eq == std::string::npos) throw std::runtime_error("--option expects key=value");
      o.lib_options.emplace_back(kv.substr(0, eq), kv.substr(eq + 1));
    } else if (a == "--number-of-tiles" || a == "-r" || a == "--resolution") {
      std::vector<unsigned> v;  // multitoken: one or two integers
      while (i + 1 < argc && is_number(argv[i + 1]) && v.size() < 2) v.push_back((unsigned)std::stoul(argv[++i]));
      if (v.empty()) throw std::runtime_error("the required argument for option '" + a + "' is missing");
      if (v.size() == 1) v.push_back(v[0]);  // ConfigParser.cpp:131-133,139-141
      (a == "--number-of-tiles" ? o.n_tiles : o.resolution) = v;
    } else if (!a.empty() && a[0] == '-')
      throw std::runtime_error("unrecognised option '" + a + "'");
    else
      o.scene_file = a;  // positional scene-file
  }
  return true;
}

// Radiance .hdr, flat (non-RLE) RGBE scanlines; top-to-bottom like stbi_write_hdr
void saveHDR(const std::string& base, const float* rgba, int w, int h) {
  std::string fn = base + ".hdr";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + fn);
  fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", h, w);
  std::vector<unsigned char> row((size_t)w * 4);
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const float* p = rgba + 4 * ((size_t)y * w + x);
      float m = std::max(p[0], std::max(p[1], p[2]));
      unsigned char* q = &row[4 * (size_t)x];
      if (!(m > 1e-32f)) {
        q[0] = q[1] = q[2] = q[3] = 0;
      } else {
        int e;
        float s = std::frexp(m, &e) * 256.0f / m;
        q[0] = (unsigned char)(p[0] * s), q[1] = (unsigned char)(p[1] * s), q[2] = (unsigned char)(p[2] * s);
        q[3] = (unsigned char)(e + 128);
      }
    }
    fwrite(row.data(), 1, row.size(), f);
  }
  fclose(f);
  std::cout << "Saved " << fn << ".\n";
}

// Image::savePNG (Image.cpp:35-56): clamp to [0,1], x255, truncate, 8-bit RGB.  Written with
// zlib directly (stored PNG chunks: signature, IHDR, one IDAT, IEND); no stb.
void savePNG(const std::string& base, const float* rgba, int w, int h) {
  std::vector<unsigned char> raw((size_t)h * (3 * (size_t)w + 1));
  for (int y = 0; y < h; ++y) {
    unsigned char* row = &raw[(size_t)y * (3 * (size_t)w + 1)];
    row[0] = 0;  // filter type: none
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) {
        float v = rgba[4 * ((size_t)y * w + x) + c];
        v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);  // NaN pixels (reference quirk) become 0
        if (!(v == v)) v = 0.f;
        row[1 + 3 * x + c] = (unsigned char)(v * 255.f);
      }
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("png: deflate failed");
  std::string fn = base + ".png";
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write " + fn);
  auto be32 = [](unsigned char* p, uint32_t v) { p[0] = v >> 24, p[1] = v >> 16, p[2] = v >> 8, p[3] = v; };
  auto chunk = [&](const char* type, const unsigned char* data, uint32_t n) {
    unsigned char hdr[8];
    be32(hdr, n);
    memcpy(hdr + 4, type, 4);
    fwrite(hdr, 1, 8, f);
    if (n) fwrite(data, 1, n, f);
    uLong c = crc32(0L, hdr + 4, 4);
    if (n) c = crc32(c, data, n);
    unsigned char crc[4];
    be32(crc, (uint32_t)c);
    fwrite(crc, 1, 4, f);
  };
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  fwrite(sig, 1, 8, f);
  unsigned char ihdr[13];
  be32(ihdr, (uint32_t)w), be32(ihdr + 4, (uint32_t)h);
  ihdr[8] = 8, ihdr[9] = 2, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;  // 8-bit, colour type 2 (RGB)
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), (uint32_t)zlen);
  chunk("IEND", nullptr, 0);
  fclose(f);
  std::cout << "Saved " << fn << ".\n";
}

int runTest(const Options& o, const Scene& scene) {
  TilingConfig tiling({o.resolution[0], o.resolution[1]}, {o.n_tiles[0], o.n_tiles[1]});
  std::vector<float> times;
  double mean_time = 0;
  const size_t npx = (size_t)o.resolution[0] * o.resolution[1];
  std::string out_name = o.output.empty() ? ("algorithm_" + o.algorithm + "_kernel_" + o.kernel + "_iter_" +
                                             std::to_string(o.iterations))
                                          : o.output;
  for (unsigned i = 0; i < o.trials; ++i) {
    printf("---------------------------------------------------------------trial : %u \n", i);
    auto t0 = std::chrono::steady_clock::now();
    float* pixels = nullptr;  // Image.cpp:7-15: pinned write-combined host image
    cuda_ck(cudaSetDevice(o.device), "cudaSetDevice");
    cuda_ck(cudaHostAlloc((void**)&pixels, npx * 16, cudaHostAllocWriteCombined), "cudaHostAlloc");
    std::fill(pixels, pixels + npx * 4, 0.f);
    // --option pairs reach the launcher through the C ABI before init() / setScene()
    std::unique_ptr<AbstractRenderer> renderer;
    if (o.gpus > 1 || o.use_group) {
      const int mode = o.shard == "tiles" ? CVR_SHARD_TILES : o.shard == "spp" ? CVR_SHARD_SPP : o.shard == "balanced" ? CVR_SHARD_BALANCED : -1;
      if (mode < 0) throw std::runtime_error("--shard expects tiles | spp | balanced");
      renderer = std::make_unique<GroupVolPath>(o.kernel, scene, tiling, o.iterations, o.gpus, mode, o.lib_options);
    } else {
      renderer = createRenderer(o.kernel, scene, tiling, o.iterations, o.device, o.lib_options);
    }
    auto t1 = std::chrono::steady_clock::now();
    printf("initialization time : %.2f sec \n", std::chrono::duration<float>(t1 - t0).count());
    Buffer2D out = make_buffer2D_float4(pixels, o.resolution[0], o.resolution[1]);
    t0 = std::chrono::steady_clock::now();
    renderer->render(out);
    t1 = std::chrono::steady_clock::now();
    float dt = std::chrono::duration<float>(t1 - t0).count();
    printf("rendering time      : %.4f sec \n", dt);
    if (i > 0) {  // discard the first iteration (Main.cpp:79-81)
      times.push_back(dt);
      mean_time += dt;
    }
    std::vector<float> copy(pixels, pixels + npx * 4);  // write-combined memory: read once, linearly
    if (!o.dump_raw.empty()) {
      FILE* f = fopen(o.dump_raw.c_str(), "wb");
      if (!f) throw std::runtime_error("cannot write " + o.dump_raw);
      fwrite(copy.data(), sizeof(float), copy.size(), f);
      fclose(f);
    }
    saveHDR(out_name, copy.data(), (int)o.resolution[0], (int)o.resolution[1]);
    if (o.png) savePNG(out_name, copy.data(), (int)o.resolution[0], (int)o.resolution[1]);
    renderer.reset();
    cudaFreeHost(pixels);
  }
  if (o.trials > 1) {
    mean_time /= times.size();
    double var = 0;
    for (float t : times) var += (t - mean_time) * (t - mean_time);
    var /= times.size();
    printf("execution mean time of %.4f sec on %zu iterations and std %.5f \n", mean_time, times.size(), std::sqrt(var));
    printf("paths per sec %lf \n", (double)o.resolution[0] * o.resolution[1] * o.iterations / mean_time);
  }
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  try {
    Options o;
    if (!parse(argc, argv, o)) return 0;
    if (o.scene_file.empty()) throw std::runtime_error("Error: no scene file provided");  // ConfigParser.cpp:71-73
    // ConfigParser.cpp:84-103: "Auto" resolves by extension (makeSceneBuilder, also behind cvr_scene_file_load)
    std::string type;
    std::unique_ptr<SceneBuilder> builder = makeSceneBuilder(o.scene_file, o.scene_type, &type);
    if (o.scene_type == "Auto" && type != "Synth") std::cout << print_prefix << "Auto-detected scene type: " << type << "\n";
    SceneAssembler assembler;
    assembler.setBuilder(std::move(builder));
    Scene scene = assembler.getScene();
    if (o.algorithm != "cudaVolPath") throw std::runtime_error("Error: algorithm '" + o.algorithm + "' unknown (cudaVolPath)");
    std::cout << print_prefix << "algorithm set to " << o.algorithm << ".\n";
    std::cout << print_prefix << "kernel set to " << o.kernel << ".\n";
    std::cout << print_prefix << "iterations set to " << o.iterations << ".\n";
    // --resolution always overrides the film size (its default makes count() == 1, Q5)
    scene.getCamera()->setResolution((int)o.resolution[0], (int)o.resolution[1]);
    if (o.dump_scene) {
      const HostMedium& m = scene.getMedium();
      auto rtv = scene.getCamera()->getRasterToView();
      double dsum = 0, asum = 0;
      for (float v : m.density_volume.data) dsum += v;
      for (float v : m.albedo_volume.data) asum += v;
      printf("scene density %u %u %u albedo %u %u %u\n", m.density_volume.nx, m.density_volume.ny, m.density_volume.nz,
             m.albedo_volume.nx, m.albedo_volume.ny, m.albedo_volume.nz);
      printf("box %.9g %.9g %.9g %.9g %.9g %.9g\n", m.density_AABB.box_min.x, m.density_AABB.box_min.y,
             m.density_AABB.box_min.z, m.density_AABB.box_max.x, m.density_AABB.box_max.y, m.density_AABB.box_max.z);
      printf("scale %.9g max_density %.9g fov %.9g rtv %.9g %.9g\n", m.scale, m.max_density,
             scene.getCamera()->getFovX(), rtv[0], rtv[1]);
      printf("sums %.9g %.9g\n", dsum, asum);
      return 0;
    }
    if (o.interactive)
      std::cout << print_prefix << "interactive view is not available in this build; rendering offline.\n";
    return runTest(o, scene);
  } catch (const std::exception& e) {
    std::cerr << print_prefix << "Error: " << e.what() << "\n";
    return 1;
  }
}
