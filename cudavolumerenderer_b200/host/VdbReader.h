// VdbReader.h -- dependency-free reader for OpenVDB .vdb files.
//
// Replaces the reference's vdb_adapter (implementation/vdb_adapter/VDBAdapter.{h,cpp}),
// which links OpenVDB + blosc + TBB, for exactly what the path needs: the FloatGrid
// "density" and the Vec3SGrid "albedo" of a file written by scripts/convert-mhd/mhd_to_vdb.py
// (tree configuration Tree_{float,vec3s}_5_4_3, file versions 222..224, compression
// "blosc + active values", "zip + active values" or none).  Written from the published
// OpenVDB file layout (io/Archive, io/Compression, tree/{Root,Internal,Leaf}Node read paths)
// and the c-blosc 1.x container + LZ4 block formats; no OpenVDB, blosc or LZ4 code is used.
//
// What is kept of a grid: its 8^3 leaves (origin, 512-bit value mask, 512 values) and its
// tiles -- i.e. the sparse form, so that a bricked device layout can be built directly from
// the leaves (8^3 leaf = one brick); densify() reproduces VDBAdapter's dense copies
// (VDBAdapter.cpp:46-114) for the dense path.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace cvrvdb {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------ byte stream
class Reader {
  const uint8_t* p_;
  size_t n_, pos_ = 0;

 public:
  Reader(const uint8_t* p, size_t n) : p_(p), n_(n) {}
  size_t pos() const { return pos_; }
  size_t size() const { return n_; }
  void seek(size_t q) {
    if (q > n_) throw Error("vdb: seek past the end of the file");
    pos_ = q;
  }
  const uint8_t* take(size_t k) {
    if (k > n_ - pos_) throw Error("vdb: truncated file");
    const uint8_t* r = p_ + pos_;
    pos_ += k;
    return r;
  }
  template <class T>
  T get() {
    T v;
    std::memcpy(&v, take(sizeof(T)), sizeof(T));
    return v;
  }
  std::string str() {  // u32 length + bytes (io::readString)
    uint32_t n = get<uint32_t>();
    if (n > (1u << 24)) throw Error("vdb: implausible string length");
    const uint8_t* s = take(n);
    return std::string(reinterpret_cast<const char*>(s), n);
  }
};

// ------------------------------------------------------------------ LZ4 block format
// Sequence = token (literal length << 4 | match length - 4), optional 255-run length bytes,
// literals, 2-byte little-endian offset, optional match length bytes.  The last sequence
// ends after its literals.
inline void lz4_block_decode(const uint8_t* src, size_t n_src, uint8_t* dst, size_t n_dst) {
  size_t ip = 0, op = 0;
  while (ip < n_src) {
    const unsigned token = src[ip++];
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= n_src) throw Error("vdb: corrupt LZ4 block (literal length)");
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (lit > n_src - ip || lit > n_dst - op) throw Error("vdb: corrupt LZ4 block (literals)");
    std::memcpy(dst + op, src + ip, lit);
    ip += lit, op += lit;
    if (ip >= n_src) break;  // last sequence: literals only
    if (n_src - ip < 2) throw Error("vdb: corrupt LZ4 block (offset)");
    const size_t off = src[ip] | (size_t(src[ip + 1]) << 8);
    ip += 2;
    size_t len = token & 15u;
    if (len == 15) {
      unsigned b;
      do {
        if (ip >= n_src) throw Error("vdb: corrupt LZ4 block (match length)");
        b = src[ip++];
        len += b;
      } while (b == 255);
    }
    len += 4;
    if (off == 0 || off > op || len > n_dst - op) throw Error("vdb: corrupt LZ4 block (match)");
    for (size_t i = 0; i < len; ++i) dst[op + i] = dst[op + i - off];  // overlap is the point
    op += len;
  }
  if (op != n_dst) throw Error("vdb: LZ4 block decoded to the wrong size");
}

// ------------------------------------------------------------------ c-blosc 1.x container
// 16-byte header {version, versionlz, flags, typesize, nbytes, blocksize, cbytes}; flags:
// 0x1 byte shuffle, 0x2 stored (memcpy), 0x4 bit shuffle, 0x10 do-not-split, bits 5-7 codec
// (0 blosclz, 1 lz4/lz4hc, 2 snappy, 3 zlib, 4 zstd).  Then one int32 start offset per block;
// a block holds `typesize` byte-plane streams when it is split, else one stream; a stream is
// {int32 compressed size, bytes} and is stored raw when that size equals its plain size.
inline void blosc_decode(const uint8_t* src, size_t n_src, uint8_t* dst, size_t n_dst) {
  if (n_src < 16) throw Error("vdb: blosc chunk shorter than its header");
  const unsigned flags = src[2], typesize = src[3];
  uint32_t nbytes, blocksize, cbytes;
  std::memcpy(&nbytes, src + 4, 4), std::memcpy(&blocksize, src + 8, 4), std::memcpy(&cbytes, src + 12, 4);
  if (nbytes != n_dst) throw Error("vdb: blosc chunk has the wrong uncompressed size");
  if (cbytes > n_src) throw Error("vdb: blosc chunk is truncated");
  if (nbytes == 0) return;
  if (flags & 0x2) {  // stored
    if (n_src < 16 + (size_t)nbytes) throw Error("vdb: stored blosc chunk is truncated");
    std::memcpy(dst, src + 16, nbytes);
    return;
  }
  if (flags & 0x4) throw Error("vdb: blosc bit-shuffle is not supported");
  if (blocksize == 0 || typesize == 0) throw Error("vdb: corrupt blosc header");
  const unsigned codec = flags >> 5;
  const bool shuffle = (flags & 0x1) && typesize > 1;
  const bool dont_split = (flags & 0x10) != 0;
  const uint32_t nblocks = (nbytes + blocksize - 1) / blocksize;
  if (16 + 4 * (size_t)nblocks > n_src) throw Error("vdb: corrupt blosc block table");
  std::vector<uint8_t> tmp(blocksize);
  for (uint32_t b = 0; b < nblocks; ++b) {
    int32_t start;
    std::memcpy(&start, src + 16 + 4 * (size_t)b, 4);
    if (start < 0 || (size_t)start > n_src) throw Error("vdb: corrupt blosc block offset");
    const uint32_t bsize = (b == nblocks - 1 && nbytes % blocksize) ? nbytes % blocksize : blocksize;
    const bool leftover = bsize != blocksize;
    const bool split = !dont_split && typesize <= 16 && (blocksize / typesize) >= 128 && !leftover;
    const uint32_t nsplits = split ? typesize : 1;
    const uint32_t neblock = bsize / nsplits;
    uint8_t* out = shuffle ? tmp.data() : dst + (size_t)b * blocksize;
    size_t ip = (size_t)start;
    for (uint32_t j = 0; j < nsplits; ++j) {
      if (ip + 4 > n_src) throw Error("vdb: corrupt blosc stream header");
      int32_t cb;
      std::memcpy(&cb, src + ip, 4);
      ip += 4;
      if (cb < 0 || ip + (size_t)cb > n_src) throw Error("vdb: corrupt blosc stream size");
      if ((uint32_t)cb == neblock) {
        std::memcpy(out + (size_t)j * neblock, src + ip, neblock);
      } else if (codec == 1) {
        lz4_block_decode(src + ip, (size_t)cb, out + (size_t)j * neblock, neblock);
      } else if (codec == 3) {
        uLongf dl = neblock;
        if (uncompress(out + (size_t)j * neblock, &dl, src + ip, (uLong)cb) != Z_OK || dl != neblock)
          throw Error("vdb: zlib stream inside blosc failed to inflate");
      } else {
        throw Error("vdb: blosc codec " + std::to_string(codec) + " is not supported (lz4 and zlib are)");
      }
      ip += (size_t)cb;
    }
    if (shuffle) {  // byte j of element i sits at tmp[j * n_elem + i]; the tail is verbatim
      uint8_t* d = dst + (size_t)b * blocksize;
      const uint32_t n_elem = bsize / typesize, rem = bsize % typesize;
      for (uint32_t i = 0; i < n_elem; ++i)
        for (uint32_t j = 0; j < typesize; ++j) d[(size_t)i * typesize + j] = tmp[(size_t)j * n_elem + i];
      std::memcpy(d + (size_t)n_elem * typesize, tmp.data() + (size_t)n_elem * typesize, rem);
    }
  }
}

// ------------------------------------------------------------------ grid model
struct Leaf {
  int32_t origin[3];
  uint64_t mask[8];           // bit n = voxel n active, n = (x&7)<<6 | (y&7)<<3 | (z&7)
  std::vector<float> values;  // 512 * channels, inactive voxels already resolved
  bool on(unsigned n) const { return (mask[n >> 6] >> (n & 63)) & 1u; }
};
struct Tile {
  int32_t origin[3];
  int32_t dim;  // edge length in voxels: 8 (level-1 entry), 128 (level-2 entry), 4096 (root tile)
  bool active;
  float value[3];
};
struct Grid {
  std::string name, type;
  int channels = 1;
  bool half = false;
  uint32_t compression = 0;
  float background[3] = {0, 0, 0};
  std::map<std::string, std::string> meta;  // decoded where the type is known
  std::vector<Leaf> leaves;
  std::vector<Tile> tiles;
  uint64_t active_voxels = 0;  // leaf voxels + the voxels covered by active tiles
  bool has_active = false;
  int32_t bbox_min[3] = {0, 0, 0}, bbox_max[3] = {-1, -1, -1};  // evalActiveVoxelBoundingBox()
  int32_t dim(int a) const { return has_active ? bbox_max[a] - bbox_min[a] + 1 : 0; }

  // VDBAdapter::get{Density,Albedo}DataAsLinearArray (VDBAdapter.cpp:57-114): a dense
  // x-fastest array over the active bounding box, `inactive` everywhere, active values at
  // their coordinate.  An active TILE contributes one value at its origin only, because the
  // reference's ValueOn iterator reports a tile once (coord = its origin).  out_channels
  // may exceed the grid's (float4 albedo: w = pad_w).
  void densify(float* out, int out_channels, const float* inactive, float pad_w = 1.0f) const {
    const size_t nx = (size_t)dim(0), ny = (size_t)dim(1), nz = (size_t)dim(2);
    for (size_t i = 0; i < nx * ny * nz; ++i)
      for (int c = 0; c < out_channels; ++c) out[i * out_channels + c] = c < channels ? inactive[c] : pad_w;
    auto put = [&](int32_t x, int32_t y, int32_t z, const float* v) {
      const size_t i = (size_t)(x - bbox_min[0]) + nx * ((size_t)(y - bbox_min[1]) + ny * (size_t)(z - bbox_min[2]));
      for (int c = 0; c < channels && c < out_channels; ++c) out[i * out_channels + c] = v[c];
    };
    for (const Leaf& L : leaves)
      for (unsigned n = 0; n < 512; ++n)
        if (L.on(n)) put(L.origin[0] + (int)(n >> 6), L.origin[1] + (int)((n >> 3) & 7), L.origin[2] + (int)(n & 7), &L.values[(size_t)n * channels]);
    for (const Tile& T : tiles)
      if (T.active) put(T.origin[0], T.origin[1], T.origin[2], T.value);
  }
};

// ------------------------------------------------------------------ file parser
class File {
 public:
  uint32_t file_version = 0, lib_major = 0, lib_minor = 0;
  std::string uuid;
  std::map<std::string, std::string> meta;
  std::vector<Grid> grids;

  const Grid* find(const std::string& name) const {
    for (const Grid& g : grids)
      if (g.name == name) return &g;
    return nullptr;
  }

  static File load(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw Error("vdb: cannot open '" + path + "'");
    const std::streamoff n = f.tellg();
    std::vector<uint8_t> buf((size_t)n);
    f.seekg(0);
    f.read(reinterpret_cast<char*>(buf.data()), n);
    if (!f) throw Error("vdb: cannot read '" + path + "'");
    return parse(buf.data(), buf.size());
  }

  static File parse(const uint8_t* data, size_t n) {
    Reader r(data, n);
    File F;
    if (r.get<int64_t>() != 0x56444220) throw Error("vdb: not an OpenVDB file (bad magic)");
    F.file_version = r.get<uint32_t>();
    if (F.file_version < 222)
      throw Error("vdb: file version " + std::to_string(F.file_version) + " predates node-mask compression (222); not supported");
    F.lib_major = r.get<uint32_t>(), F.lib_minor = r.get<uint32_t>();
    const bool has_offsets = r.get<uint8_t>() != 0;
    F.uuid = std::string(reinterpret_cast<const char*>(r.take(36)), 36);
    read_meta(r, F.meta);
    const uint32_t n_grids = r.get<uint32_t>();
    if (!has_offsets && n_grids > 1) throw Error("vdb: files without grid offsets (streamed) are not supported");
    for (uint32_t i = 0; i < n_grids; ++i) {
      Grid g;
      g.name = r.str();
      // the unique name may carry a 0x1e-separated numeric suffix (GridDescriptor::nameAsString)
      if (size_t s = g.name.find('\x1e'); s != std::string::npos) g.name.resize(s);
      g.type = r.str();
      static const std::string kHalf = "_HalfFloat";
      if (g.type.size() > kHalf.size() && g.type.compare(g.type.size() - kHalf.size(), kHalf.size(), kHalf) == 0) {
        g.half = true;
        g.type.resize(g.type.size() - kHalf.size());
      }
      const std::string parent = r.str();
      const int64_t grid_pos = r.get<int64_t>(), block_pos = r.get<int64_t>(), end_pos = r.get<int64_t>();
      if (!parent.empty()) throw Error("vdb: instanced grids are not supported ('" + g.name + "')");
      if (has_offsets) r.seek((size_t)grid_pos);
      if (g.type == "Tree_float_5_4_3")
        g.channels = 1;
      else if (g.type == "Tree_vec3s_5_4_3")
        g.channels = 3;
      else {  // not a grid the path uses: remember it, skip its payload
        if (!has_offsets) throw Error("vdb: cannot skip grid type '" + g.type + "' without offsets");
        g.channels = 0;
        F.grids.push_back(std::move(g));
        r.seek((size_t)end_pos);
        continue;
      }
      g.compression = r.get<uint32_t>();  // per-grid since version 222
      read_meta(r, g.meta);
      if (auto it = g.meta.find("is_saved_as_half_float"); it != g.meta.end() && it->second == "1") g.half = true;
      skip_transform(r);
      read_tree(r, g);
      if (has_offsets && r.pos() != (size_t)block_pos) throw Error("vdb: topology of '" + g.name + "' does not end at its block offset");
      read_buffers(r, g);
      if (has_offsets && r.pos() != (size_t)end_pos) throw Error("vdb: buffers of '" + g.name + "' do not end at the grid's end offset");
      finish(g);
      F.grids.push_back(std::move(g));
    }
    return F;
  }

 private:
  enum : uint32_t { COMPRESS_ZIP = 1, COMPRESS_ACTIVE_MASK = 2, COMPRESS_BLOSC = 4 };

  static void read_meta(Reader& r, std::map<std::string, std::string>& out) {
    const uint32_t n = r.get<uint32_t>();
    for (uint32_t i = 0; i < n; ++i) {
      const std::string name = r.str(), type = r.str();
      const uint32_t size = r.get<uint32_t>();
      const uint8_t* p = r.take(size);
      auto num = [&](auto tag, int count) {
        using T = decltype(tag);
        std::string s;
        for (int k = 0; k < count && (size_t)(k + 1) * sizeof(T) <= size; ++k) {
          T v;
          std::memcpy(&v, p + k * sizeof(T), sizeof(T));
          s += (k ? " " : "") + std::to_string(v);
        }
        return s;
      };
      if (type == "string")
        out[name] = std::string(reinterpret_cast<const char*>(p), size);
      else if (type == "bool")
        out[name] = size && p[0] ? "1" : "0";
      else if (type == "int32")
        out[name] = num(int32_t(), 1);
      else if (type == "int64")
        out[name] = num(int64_t(), 1);
      else if (type == "float")
        out[name] = num(float(), 1);
      else if (type == "double")
        out[name] = num(double(), 1);
      else if (type == "vec3i")
        out[name] = num(int32_t(), 3);
      else if (type == "vec3s")
        out[name] = num(float(), 3);
      else if (type == "vec3d")
        out[name] = num(double(), 3);
      // anything else (e.g. "__delayedload") is size-prefixed and simply skipped
    }
  }

  // math::Transform::read: map type name + the map's own fields.  The reference reads the
  // world box and then ignores it (VDBSceneBuilder.h:70-77, Q4), so only the size matters.
  static void skip_transform(Reader& r) {
    const std::string t = r.str();
    size_t bytes;
    if (t == "UniformScaleMap" || t == "ScaleMap")
      bytes = 5 * 24;
    else if (t == "UniformScaleTranslateMap" || t == "ScaleTranslateMap")
      bytes = 6 * 24;
    else if (t == "TranslationMap")
      bytes = 24;
    else if (t == "AffineMap" || t == "UnitaryMap")
      bytes = 128;
    else
      throw Error("vdb: transform map '" + t + "' is not supported");
    r.take(bytes);
  }

  static float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0) {
      if (m == 0)
        u = s;
      else {  // subnormal
        int sh = 0;
        uint32_t mm = m;
        while (!(mm & 1024u)) mm <<= 1, ++sh;
        u = s | ((uint32_t)(127 - 15 - sh + 1) << 23) | ((mm & 1023u) << 13);
      }
    } else if (e == 31)
      u = s | 0x7f800000u | (m << 13);
    else
      u = s | ((e - 15 + 127) << 23) | (m << 13);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }

  // io::readData: `count` values of `channels` floats (or halfs), through the grid's codec
  static void read_data(Reader& r, const Grid& g, size_t count, float* out) {
    const size_t elem = (g.half ? 2 : 4) * (size_t)g.channels, bytes = count * elem;
    std::vector<uint8_t> raw(bytes);
    if (g.compression & COMPRESS_BLOSC) {
      const int64_t nz = r.get<int64_t>();
      if (nz <= 0) {  // stored uncompressed, -nz bytes
        if ((uint64_t)(-nz) != bytes) throw Error("vdb: uncompressed chunk has the wrong size");
        std::memcpy(raw.data(), r.take(bytes), bytes);
      } else {
        blosc_decode(r.take((size_t)nz), (size_t)nz, raw.data(), bytes);
      }
    } else if (g.compression & COMPRESS_ZIP) {
      const int64_t nz = r.get<int64_t>();
      if (nz <= 0) {
        if ((uint64_t)(-nz) != bytes) throw Error("vdb: uncompressed chunk has the wrong size");
        std::memcpy(raw.data(), r.take(bytes), bytes);
      } else {
        uLongf dl = (uLongf)bytes;
        if (uncompress(raw.data(), &dl, r.take((size_t)nz), (uLong)nz) != Z_OK || dl != bytes)
          throw Error("vdb: zip chunk failed to inflate");
      }
    } else {
      std::memcpy(raw.data(), r.take(bytes), bytes);
    }
    const size_t nf = count * (size_t)g.channels;
    if (g.half) {
      for (size_t i = 0; i < nf; ++i) {
        uint16_t h;
        std::memcpy(&h, raw.data() + 2 * i, 2);
        out[i] = half_to_float(h);
      }
    } else {
      std::memcpy(out, raw.data(), nf * 4);
    }
  }

  // one ValueType at full precision (backgrounds, tiles and per-node inactive values are
  // never stored as half)
  static void read_value(Reader& r, const Grid& g, float* v) {
    for (int c = 0; c < g.channels; ++c) v[c] = r.get<float>();
  }

  // io::readCompressedValues: `count` values of a node whose value mask is `mask`
  // (count/64 words).  With active-mask compression only the ACTIVE values are stored and
  // the inactive ones are rebuilt from a per-node code: 0 all +background, 1 all -background,
  // 2 one other value, 3 selection mask between -/+ background, 4 mask between background
  // and one value, 5 mask between two values, 6 everything stored.
  static void read_compressed(Reader& r, const Grid& g, size_t count, const uint64_t* mask, float* out) {
    const int C = g.channels;
    const int8_t code = r.get<int8_t>();
    if (code < 0 || code > 6) throw Error("vdb: corrupt node compression code");
    float inactive0[3], inactive1[3];
    for (int c = 0; c < 3; ++c) inactive0[c] = g.background[c], inactive1[c] = g.background[c];
    if (code == 1)
      for (int c = 0; c < C; ++c) inactive0[c] = -g.background[c];
    if (code == 3)
      for (int c = 0; c < C; ++c) inactive0[c] = -g.background[c];  // mask OFF -> -background, ON -> +background
    if (code == 2 || code == 4 || code == 5) {
      read_value(r, g, inactive0);
      if (code == 5) read_value(r, g, inactive1);
    }
    std::vector<uint64_t> sel;
    if (code == 3 || code == 4 || code == 5) {
      sel.resize(count / 64);
      std::memcpy(sel.data(), r.take(count / 8), count / 8);
    }
    const bool mask_compressed = (g.compression & COMPRESS_ACTIVE_MASK) && code != 6;
    if (!mask_compressed) {
      read_data(r, g, count, out);
      return;
    }
    size_t n_on = 0;
    for (size_t w = 0; w < count / 64; ++w) n_on += (size_t)__builtin_popcountll(mask[w]);
    std::vector<float> act(n_on * (size_t)C);
    read_data(r, g, n_on, act.data());
    size_t k = 0;
    for (size_t i = 0; i < count; ++i) {
      const bool on = (mask[i >> 6] >> (i & 63)) & 1u;
      const float* src;
      if (on)
        src = &act[(k++) * (size_t)C];
      else if (!sel.empty() && ((sel[i >> 6] >> (i & 63)) & 1u))
        src = inactive1;
      else
        src = inactive0;
      for (int c = 0; c < C; ++c) out[i * (size_t)C + c] = src[c];
    }
  }

  struct PendingLeaf {  // topology pass: origin + mask; the buffer pass fills the values in order
    int32_t origin[3];
  };

  // InternalNode<LOG2>::readTopology with children of edge `child_dim` voxels
  static void read_internal(Reader& r, Grid& g, const int32_t origin[3], int log2dim, int child_dim, int depth) {
    const size_t n = (size_t)1 << (3 * log2dim);
    std::vector<uint64_t> child(n / 64), value(n / 64);
    std::memcpy(child.data(), r.take(n / 8), n / 8);
    std::memcpy(value.data(), r.take(n / 8), n / 8);
    std::vector<float> vals(n * (size_t)g.channels);
    read_compressed(r, g, n, value.data(), vals.data());
    const int mask_dim = (1 << log2dim) - 1;
    for (size_t i = 0; i < n; ++i) {
      const bool is_child = (child[i >> 6] >> (i & 63)) & 1u;
      int32_t o[3] = {origin[0] + (int32_t)((i >> (2 * log2dim)) & mask_dim) * child_dim,
                      origin[1] + (int32_t)((i >> log2dim) & mask_dim) * child_dim,
                      origin[2] + (int32_t)(i & mask_dim) * child_dim};
      if (is_child) {
        if (depth == 0) {
          read_internal(r, g, o, 4, 8, 1);
        } else {  // LeafNode::readTopology: the 512-bit value mask
          Leaf L;
          L.origin[0] = o[0], L.origin[1] = o[1], L.origin[2] = o[2];
          std::memcpy(L.mask, r.take(64), 64);
          g.leaves.push_back(std::move(L));
        }
      } else if ((value[i >> 6] >> (i & 63)) & 1u) {  // active tile
        Tile T;
        T.origin[0] = o[0], T.origin[1] = o[1], T.origin[2] = o[2];
        T.dim = child_dim, T.active = true;
        for (int c = 0; c < 3; ++c) T.value[c] = c < g.channels ? vals[i * (size_t)g.channels + c] : 0.f;
        g.tiles.push_back(T);
      }
    }
  }

  // TreeBase::readTopology + RootNode::readTopology
  static void read_tree(Reader& r, Grid& g) {
    const int32_t buffer_count = r.get<int32_t>();
    if (buffer_count != 1) throw Error("vdb: multi-buffer trees are not supported");
    // the background is stored at full precision even for half grids
    for (int c = 0; c < g.channels; ++c) g.background[c] = r.get<float>();
    const uint32_t n_tiles = r.get<uint32_t>(), n_children = r.get<uint32_t>();
    for (uint32_t i = 0; i < n_tiles; ++i) {
      Tile T;
      for (int a = 0; a < 3; ++a) T.origin[a] = r.get<int32_t>();
      for (int c = 0; c < 3; ++c) T.value[c] = c < g.channels ? r.get<float>() : 0.f;
      T.active = r.get<uint8_t>() != 0;
      T.dim = 4096;
      if (T.active) g.tiles.push_back(T);
    }
    for (uint32_t i = 0; i < n_children; ++i) {
      int32_t o[3];
      for (int a = 0; a < 3; ++a) o[a] = r.get<int32_t>();
      read_internal(r, g, o, 5, 128, 0);
    }
  }

  // Tree::readBuffers: the leaves in topology order; each = value mask again + values
  static void read_buffers(Reader& r, Grid& g) {
    for (Leaf& L : g.leaves) {
      std::memcpy(L.mask, r.take(64), 64);
      L.values.resize(512 * (size_t)g.channels);
      read_compressed(r, g, 512, L.mask, L.values.data());
    }
  }

  static void finish(Grid& g) {  // evalActiveVoxelBoundingBox + active voxel count
    auto grow = [&](const int32_t lo[3], const int32_t hi[3]) {
      for (int a = 0; a < 3; ++a) {
        if (!g.has_active || lo[a] < g.bbox_min[a]) g.bbox_min[a] = lo[a];
        if (!g.has_active || hi[a] > g.bbox_max[a]) g.bbox_max[a] = hi[a];
      }
      g.has_active = true;
    };
    for (const Leaf& L : g.leaves) {
      int32_t lo[3] = {8, 8, 8}, hi[3] = {-1, -1, -1};
      size_t n_on = 0;
      for (unsigned n = 0; n < 512; ++n)
        if (L.on(n)) {
          const int32_t c[3] = {(int32_t)(n >> 6), (int32_t)((n >> 3) & 7), (int32_t)(n & 7)};
          for (int a = 0; a < 3; ++a) lo[a] = c[a] < lo[a] ? c[a] : lo[a], hi[a] = c[a] > hi[a] ? c[a] : hi[a];
          ++n_on;
        }
      if (!n_on) continue;
      g.active_voxels += n_on;
      for (int a = 0; a < 3; ++a) lo[a] += L.origin[a], hi[a] += L.origin[a];
      grow(lo, hi);
    }
    for (const Tile& T : g.tiles) {
      if (!T.active) continue;
      const int32_t hi[3] = {T.origin[0] + T.dim - 1, T.origin[1] + T.dim - 1, T.origin[2] + T.dim - 1};
      grow(T.origin, hi);
      g.active_voxels += (uint64_t)T.dim * T.dim * T.dim;
    }
  }
};

}  // namespace cvrvdb
