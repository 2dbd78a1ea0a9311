// vi_hdf5.cu -- a minimal native HDF5 reader for the data sets the reference's test program feeds IndexBuilder with
// (VectorIndex.MainTest/Program.cs:183-260: GetHdf5DatasetSize / GetHdf5Dataset over `/train` and `/test` of an
// ANN-benchmarks file such as deep-image-96-angular.hdf5, through HDF5-CSharp 1.19.0 = libhdf5).  Neither libhdf5 nor
// h5py exists in this image, so the container format is parsed here, from the published HDF5 File Format
// Specification (version 3.0), for exactly what those files contain:
//   superblock version 0/1 (libver "earliest", what h5py writes by default) or 2/3;
//   groups as symbol tables (B-tree v1 + local heap + SNOD nodes) or as compact link messages;
//   object headers version 1 or 2, with continuation blocks;
//   a CONTIGUOUS, unfiltered, little-endian, fixed-point or IEEE floating-point data set of rank 1 or 2.
// Chunked / compressed / dense-link-storage files are refused with a message, not mis-read.
// Host code only: it resolves (rows, cols, element type, byte offset of the data in the file); the rows then stream to
// the device through the same double-buffered pinned pipeline as the FileRangeStore records (vi_table.cu).
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "vi_common.cuh"

namespace
{
constexpr uint64_t H5_UNDEF = ~0ull;

struct H5Dataset
{
  int rank = 0;
  uint64_t dims[4] = {0, 0, 0, 0};
  int type_class = -1;  // 0 fixed point, 1 floating point
  int type_size = 0;
  bool little_endian = true;
  bool is_signed = true;
  int layout_class = -1;  // 0 compact, 1 contiguous, 2 chunked
  uint64_t data_addr = H5_UNDEF;  // absolute file offset
  uint64_t data_size = 0;
  bool filtered = false;
};

struct H5Reader
{
  int fd = -1;
  uint64_t file_size = 0;
  uint64_t base = 0;  // base address: every address in the file is relative to it
  int so = 8, sl = 8;  // size of offsets / lengths
  uint64_t root_ohdr = H5_UNDEF;
  std::string err;

  ~H5Reader()
  {
    if (fd >= 0) close(fd);
  }

  bool fail(const std::string& m)
  {
    if (err.empty()) err = m;
    return false;
  }

  bool rd(uint64_t off, void* dst, size_t n)
  {
    if (off > file_size || n > file_size - off) return fail("HDF5: read past the end of the file (truncated or not HDF5)");
    size_t done = 0;
    while (done < n)
    {
      const ssize_t g = pread(fd, (char*)dst + done, n - done, (off_t)(off + done));
      if (g <= 0) return fail(std::string("HDF5: read error: ") + strerror(errno));
      done += (size_t)g;
    }
    return true;
  }

  static uint64_t le(const unsigned char* p, int n)
  {
    uint64_t v = 0;
    for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
    return v;
  }

  // an address field: all ones = undefined
  uint64_t addr(const unsigned char* p) const
  {
    const uint64_t v = le(p, so);
    const uint64_t undef = so == 8 ? ~0ull : ((1ull << (8 * so)) - 1);
    return v == undef ? H5_UNDEF : v + base;
  }

  bool open(const char* path)
  {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) return fail("fstat failed");
    file_size = (uint64_t)st.st_size;
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    uint64_t sb = H5_UNDEF;
    for (uint64_t off = 0; off + 8 <= file_size && off <= (1ull << 24); off = off ? off * 2 : 512)
    {
      unsigned char b[8];
      if (!rd(off, b, 8)) return false;
      if (memcmp(b, sig, 8) == 0)
      {
        sb = off;
        break;
      }
    }
    if (sb == H5_UNDEF) return fail("not an HDF5 file (no superblock signature)");
    unsigned char h[128];
    const size_t have = (size_t)std::min<uint64_t>(sizeof(h), file_size - sb);
    memset(h, 0, sizeof(h));
    if (!rd(sb, h, have)) return false;
    const int version = h[8];
    if (version == 0 || version == 1)
    {
      so = h[13];
      sl = h[14];
      if ((so != 2 && so != 4 && so != 8) || (sl != 2 && sl != 4 && sl != 8)) return fail("HDF5: unsupported offset / length size");
      size_t p = 24 + (version == 1 ? 4 : 0);
      base = le(h + p, so);
      p += 4 * (size_t)so;  // base, free-space info, end of file, driver info
      // root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
      p += so;
      root_ohdr = addr(h + p);
    }
    else if (version == 2 || version == 3)
    {
      so = h[9];
      sl = h[10];
      if ((so != 2 && so != 4 && so != 8) || (sl != 2 && sl != 4 && sl != 8)) return fail("HDF5: unsupported offset / length size");
      size_t p = 12;
      base = le(h + p, so);
      p += 3 * (size_t)so;  // base, superblock extension, end of file
      root_ohdr = addr(h + p);
    }
    else
      return fail("HDF5: unsupported superblock version " + std::to_string(version));
    if (root_ohdr == H5_UNDEF) return fail("HDF5: no root group");
    return true;
  }

  struct Msg
  {
    int type;
    std::vector<unsigned char> data;
  };

  // every header message of the object at `a` (object header version 1 or 2, continuation blocks followed)
  bool messages(uint64_t a, std::vector<Msg>& out)
  {
    unsigned char h[64];
    memset(h, 0, sizeof(h));
    if (!rd(a, h, (size_t)std::min<uint64_t>(sizeof(h), file_size - a))) return false;
    struct Chunk
    {
      uint64_t off, len;
    };
    std::vector<Chunk> chunks;
    bool v2 = false;
    int hflags = 0;
    uint32_t nmsgs = 0xffffffffu;
    if (memcmp(h, "OHDR", 4) == 0)
    {
      if (h[4] != 2) return fail("HDF5: unsupported object header version");
      v2 = true;
      hflags = h[5];
      size_t p = 6;
      if (hflags & 0x20) p += 16;
      if (hflags & 0x10) p += 4;
      const int cs = 1 << (hflags & 3);
      const uint64_t c0 = le(h + p, cs);
      p += cs;
      chunks.push_back({a + p, c0});
    }
    else if (h[0] == 1)
    {
      nmsgs = (uint32_t)le(h + 2, 2);
      const uint64_t hs = le(h + 8, 4);
      chunks.push_back({a + 16, hs});  // 12-byte prefix padded to 8
    }
    else
      return fail("HDF5: not an object header");
    for (size_t ci = 0; ci < chunks.size() && out.size() < nmsgs; ++ci)
    {
      if (chunks.size() > 4096) return fail("HDF5: object header continuation loop");
      const Chunk c = chunks[ci];
      if (c.len > (64u << 20)) return fail("HDF5: unreasonable object header chunk");
      std::vector<unsigned char> buf((size_t)c.len);
      if (c.len && !rd(c.off, buf.data(), (size_t)c.len)) return false;
      size_t p = 0;
      const size_t mh = v2 ? (size_t)(4 + ((hflags & 0x04) ? 2 : 0)) : 8;
      while (p + mh <= buf.size() && out.size() < nmsgs)
      {
        int type;
        size_t size;
        if (v2)
        {
          type = buf[p];
          size = (size_t)le(&buf[p + 1], 2);
        }
        else
        {
          type = (int)le(&buf[p], 2);
          size = (size_t)le(&buf[p + 2], 2);
        }
        p += mh;
        if (p + size > buf.size()) break;  // v2: the gap before the checksum
        Msg m;
        m.type = type;
        m.data.assign(buf.begin() + (long)p, buf.begin() + (long)(p + size));
        p += size;
        if (!v2) p = (p + 7) & ~(size_t)7;
        if (type == 0x10)  // continuation: offset, length
        {
          if (m.data.size() < (size_t)(so + sl)) return fail("HDF5: short continuation message");
          uint64_t off = addr(m.data.data());
          uint64_t len = le(m.data.data() + so, sl);
          if (off == H5_UNDEF) return fail("HDF5: undefined continuation address");
          if (v2)
          {
            unsigned char sg[4];
            if (!rd(off, sg, 4)) return false;
            if (memcmp(sg, "OCHK", 4) != 0 || len < 8) return fail("HDF5: bad continuation block");
            off += 4;
            len -= 8;  // signature and checksum
          }
          chunks.push_back({off, len});
        }
        out.push_back(std::move(m));
      }
    }
    return true;
  }

  std::string heap_name(uint64_t heap, uint64_t name_off)
  {
    unsigned char h[64];
    memset(h, 0, sizeof(h));
    if (!rd(heap, h, (size_t)std::min<uint64_t>(sizeof(h), file_size - heap))) return std::string();
    if (memcmp(h, "HEAP", 4) != 0)
    {
      fail("HDF5: bad local heap");
      return std::string();
    }
    const uint64_t seg_size = le(h + 8, sl);
    const uint64_t seg = addr(h + 8 + 2 * sl);
    if (seg == H5_UNDEF || name_off >= seg_size)
    {
      fail("HDF5: bad link name offset");
      return std::string();
    }
    std::string s;
    for (uint64_t i = name_off; i < seg_size && s.size() < 1024; ++i)
    {
      char c;
      if (!rd(seg + i, &c, 1)) return std::string();
      if (!c) break;
      s.push_back(c);
    }
    return s;
  }

  // symbol-table group: depth-first over the B-tree; returns the object header address of `name`
  bool btree_find(uint64_t node, uint64_t heap, const std::string& name, uint64_t& found, int depth)
  {
    if (depth > 32) return fail("HDF5: group B-tree too deep");
    unsigned char h[8];
    if (!rd(node, h, 8)) return false;
    if (memcmp(h, "TREE", 4) == 0)
    {
      if (h[4] != 0) return fail("HDF5: not a group B-tree");
      const int level = h[5];
      const uint32_t used = (uint32_t)le(h + 6, 2);
      const uint64_t p0 = node + 8 + 2 * (uint64_t)so;  // behind the sibling addresses
      for (uint32_t i = 0; i < used && found == H5_UNDEF; ++i)
      {
        unsigned char cb[8];
        if (!rd(p0 + (uint64_t)sl + (uint64_t)i * (uint64_t)(sl + so), cb, (size_t)so)) return false;  // key0, then (child, key)*
        const uint64_t child = addr(cb);
        if (child == H5_UNDEF) continue;
        if (!btree_find(child, heap, name, found, depth + 1 + (level ? 0 : 0))) return false;
      }
      return true;
    }
    if (memcmp(h, "SNOD", 4) == 0)
    {
      const uint32_t nsym = (uint32_t)le(h + 6, 2);
      const size_t esz = 2 * (size_t)so + 24;
      std::vector<unsigned char> e(esz * nsym);
      if (nsym && !rd(node + 8, e.data(), e.size())) return false;
      for (uint32_t i = 0; i < nsym; ++i)
      {
        const unsigned char* p = e.data() + i * esz;
        const std::string nm = heap_name(heap, le(p, so));
        if (!err.empty()) return false;
        if (nm == name)
        {
          found = addr(p + so);
          return true;
        }
      }
      return true;
    }
    return fail("HDF5: bad group node");
  }

  bool child(uint64_t group, const std::string& name, uint64_t& out)
  {
    std::vector<Msg> ms;
    if (!messages(group, ms)) return false;
    out = H5_UNDEF;
    bool dense = false;
    for (const Msg& m : ms)
    {
      if (m.type == 0x11)  // symbol table: B-tree address, local heap address
      {
        if (m.data.size() < 2 * (size_t)so) return fail("HDF5: short symbol table message");
        const uint64_t bt = addr(m.data.data()), hp = addr(m.data.data() + so);
        if (bt == H5_UNDEF || hp == H5_UNDEF) return fail("HDF5: group without a symbol table");
        if (!btree_find(bt, hp, name, out, 0)) return false;
      }
      else if (m.type == 0x06)  // link message (compact storage of a new-style group)
      {
        const unsigned char* d = m.data.data();
        const size_t n = m.data.size();
        if (n < 2 || d[0] != 1) continue;
        const int fl = d[1];
        size_t p = 2;
        int ltype = 0;
        if (fl & 0x08) ltype = d[p++];
        if (fl & 0x04) p += 8;
        if (fl & 0x10) p += 1;
        const int ls = 1 << (fl & 3);
        if (p + (size_t)ls > n) continue;
        const uint64_t len = le(d + p, ls);
        p += (size_t)ls;
        if (p + len > n) continue;
        const std::string nm((const char*)d + p, (size_t)len);
        p += (size_t)len;
        if (nm == name && ltype == 0 && p + (size_t)so <= n) out = addr(d + p);
      }
      else if (m.type == 0x02)  // link info: a fractal heap address means dense storage
      {
        const unsigned char* d = m.data.data();
        size_t p = 2;
        if (m.data.size() >= 2 && (d[1] & 1)) p += 8;
        if (m.data.size() >= p + (size_t)so && addr(d + p) != H5_UNDEF) dense = true;
      }
      if (out != H5_UNDEF) return true;
    }
    if (dense) return fail("HDF5: group uses dense link storage (fractal heap), which this reader does not parse");
    return fail("HDF5: no object named '" + name + "'");
  }

  bool resolve(const char* path, uint64_t& obj)
  {
    obj = root_ohdr;
    std::string s(path ? path : "");
    size_t i = 0;
    while (i < s.size())
    {
      while (i < s.size() && s[i] == '/') ++i;
      size_t j = i;
      while (j < s.size() && s[j] != '/') ++j;
      if (j > i)
      {
        uint64_t next;
        if (!child(obj, s.substr(i, j - i), next)) return false;
        obj = next;
      }
      i = j;
    }
    return true;
  }

  bool dataset(uint64_t obj, H5Dataset& ds)
  {
    std::vector<Msg> ms;
    if (!messages(obj, ms)) return false;
    for (const Msg& m : ms)
    {
      const unsigned char* d = m.data.data();
      const size_t n = m.data.size();
      if (m.type == 0x01 && n >= 4)  // dataspace
      {
        const int ver = d[0];
        ds.rank = d[1];
        size_t p = ver == 1 ? 8 : 4;
        if (ver != 1 && ver != 2) return fail("HDF5: unsupported dataspace version");
        if (ds.rank > 4) return fail("Invalid rank.");  // Program.cs:203-206
        if (p + (size_t)ds.rank * (size_t)sl > n) return fail("HDF5: short dataspace message");
        for (int i = 0; i < ds.rank; ++i) ds.dims[i] = le(d + p + (size_t)i * (size_t)sl, sl);
      }
      else if (m.type == 0x03 && n >= 8)  // datatype
      {
        ds.type_class = d[0] & 0x0f;
        ds.little_endian = (d[1] & 1) == 0;
        ds.is_signed = (d[1] & 0x08) != 0;
        ds.type_size = (int)le(d + 4, 4);
      }
      else if (m.type == 0x08 && n >= 2)  // data layout
      {
        const int ver = d[0];
        if (ver == 3 || ver == 4)
        {
          ds.layout_class = d[1];
          if (ds.layout_class == 1)
          {
            if (n < 2 + (size_t)so + (size_t)sl) return fail("HDF5: short layout message");
            ds.data_addr = addr(d + 2);
            ds.data_size = le(d + 2 + so, sl);
          }
        }
        else if (ver == 1 || ver == 2)
        {
          const int dimn = d[1];
          ds.layout_class = d[2];
          if (ds.layout_class == 1)
          {
            if (n < 8 + (size_t)so + 4 * (size_t)dimn) return fail("HDF5: short layout message");
            ds.data_addr = addr(d + 8);
            ds.data_size = 0;  // sizes follow as 32-bit dimension sizes; taken from the dataspace instead
          }
        }
        else
          return fail("HDF5: unsupported data layout version");
      }
      else if (m.type == 0x0b)
        ds.filtered = true;
    }
    if (ds.rank == 0 && ds.type_class < 0) return fail("HDF5: not a data set");
    return true;
  }
};

struct H5Info
{
  int64_t rows = 0, cols = 0;
  int elem_class = 0, elem_bytes = 0;
  int64_t offset = 0;
};

bool h5_info(const char* path, const char* name, H5Info& out, std::string& err)
{
  H5Reader r;
  H5Dataset ds;
  uint64_t obj = 0;
  if (!r.open(path) || !r.resolve(name, obj) || !r.dataset(obj, ds))
  {
    err = r.err;
    return false;
  }
  if (ds.rank != 1 && ds.rank != 2)
  {
    err = "Invalid rank.";  // Program.cs:203-206 InvalidOperationException("Invalid rank.")
    return false;
  }
  if (ds.filtered || ds.layout_class == 2)
  {
    err = "HDF5: chunked / filtered (compressed) data sets are not supported; rewrite the data set contiguous";
    return false;
  }
  if (ds.layout_class != 1)
  {
    err = "HDF5: unsupported data layout (only contiguous data sets)";
    return false;
  }
  if ((ds.type_class != 0 && ds.type_class != 1) || !ds.little_endian ||
      (ds.type_size != 1 && ds.type_size != 2 && ds.type_size != 4 && ds.type_size != 8))
  {
    err = "HDF5: unsupported element type (little-endian integers and IEEE floats only)";
    return false;
  }
  out.rows = (int64_t)ds.dims[0];
  out.cols = ds.rank == 2 ? (int64_t)ds.dims[1] : 1;
  out.elem_class = ds.type_class;
  out.elem_bytes = ds.type_size;
  const uint64_t bytes = (uint64_t)out.rows * (uint64_t)out.cols * (uint64_t)ds.type_size;
  if (bytes == 0)
  {
    out.offset = 0;
    return true;
  }
  if (ds.data_addr == H5_UNDEF)
  {
    err = "HDF5: the data set has no storage allocated (never written)";
    return false;
  }
  if (ds.data_addr > r.file_size || bytes > r.file_size - ds.data_addr)
  {
    err = "HDF5: the data set's storage lies outside the file";
    return false;
  }
  out.offset = (int64_t)ds.data_addr;
  return true;
}
}  // namespace

// vi_table.cu: rows (no interleaved ids) streamed from `path` at `offset_bytes`; ids are first_id, first_id + 1, ...
int vi_points_add_rows_file_impl(vi_ctx* ctx, const char* path, int64_t offset_bytes, int64_t n, int64_t first_id,
                                 double* read_ms, double* total_ms);

// text of the last failure of a call made without a context (per host thread)
static thread_local std::string g_h5_err;
static int h5_fail(vi_ctx* ctx, int code, const std::string& msg)
{
  g_h5_err = msg;
  return ctx ? ctx->fail(code, msg) : code;
}
extern "C" const char* vi_hdf5_last_error(void) { return g_h5_err.c_str(); }

extern "C" int vi_hdf5_dataset_info(vi_ctx* ctx, const char* path, const char* dataset, int64_t* rows, int64_t* cols,
                                    int32_t* elem_class, int32_t* elem_bytes, int64_t* data_offset)
{
  if (!path || !dataset) return h5_fail(ctx, VI_ERR_INVALID_ARG, "null path or data set name");
  H5Info inf;
  std::string err;
  if (!h5_info(path, dataset, inf, err)) return h5_fail(ctx, VI_ERR_INVALID_ARG, err);
  if (rows) *rows = inf.rows;
  if (cols) *cols = inf.cols;
  if (elem_class) *elem_class = inf.elem_class;
  if (elem_bytes) *elem_bytes = inf.elem_bytes;
  if (data_offset) *data_offset = inf.offset;
  return VI_OK;
}

extern "C" int vi_hdf5_read_rows(vi_ctx* ctx, const char* path, const char* dataset, int64_t first_row, int64_t n, void* out,
                                 int64_t out_bytes)
{
  if (!path || !dataset || first_row < 0 || n < 0 || (n > 0 && !out))
    return h5_fail(ctx, VI_ERR_INVALID_ARG, "bad arguments");
  H5Info inf;
  std::string err;
  if (!h5_info(path, dataset, inf, err)) return h5_fail(ctx, VI_ERR_INVALID_ARG, err);
  if (first_row + n > inf.rows) return h5_fail(ctx, VI_ERR_INVALID_ARG, "rows outside the data set");
  const int64_t row_bytes = inf.cols * inf.elem_bytes;
  if (out_bytes < n * row_bytes) return h5_fail(ctx, VI_ERR_CAPACITY, "output buffer too small");
  if (n == 0) return VI_OK;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return h5_fail(ctx, VI_ERR_INVALID_ARG, std::string("cannot open ") + path);
  size_t done = 0;
  const size_t want = (size_t)(n * row_bytes);
  const off_t off = (off_t)(inf.offset + first_row * row_bytes);
  while (done < want)
  {
    const ssize_t g = pread(fd, (char*)out + done, want - done, off + (off_t)done);
    if (g <= 0)
    {
      close(fd);
      return h5_fail(ctx, VI_ERR_INVALID_ARG, "short read from the HDF5 file");
    }
    done += (size_t)g;
  }
  close(fd);
  return VI_OK;
}

extern "C" int vi_points_add_hdf5(vi_ctx* ctx, const char* path, const char* dataset, int64_t first_row, int64_t n,
                                  int64_t first_id, double* read_ms, double* total_ms)
{
  if (!ctx) return VI_ERR_INVALID_ARG;
  if (ctx->dims == 0) return ctx->fail(VI_ERR_STATE, "vi_points_reserve must be called first");
  if (!path || !dataset || first_row < 0) return ctx->fail(VI_ERR_INVALID_ARG, "bad arguments");
  H5Info inf;
  std::string err;
  if (!h5_info(path, dataset, inf, err)) return ctx->fail(VI_ERR_INVALID_ARG, err);
  if (inf.elem_class != 1 || inf.elem_bytes != 4)
    return ctx->fail(VI_ERR_INVALID_ARG, "HDF5: the vectors must be float32 (Hdf5.ReadDataset<float>, Program.cs:235)");
  if (inf.cols != ctx->dims) return ctx->fail(VI_ERR_INVALID_ARG, "Invalid length of vector.");  // FileRangeStore.cs:59-64
  if (n < 0) n = inf.rows > first_row ? inf.rows - first_row : 0;
  if (first_row + n > inf.rows) return ctx->fail(VI_ERR_INVALID_ARG, "rows outside the data set");
  VI_CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->built = false;
  ctx->pending_nq = -1;
  return vi_points_add_rows_file_impl(ctx, path, inf.offset + first_row * inf.cols * 4, n, first_id, read_ms, total_ms);
}
