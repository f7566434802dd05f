// Uncompressed TIFF / BigTIFF page reads straight into caller-owned (pinned) host buffers.
//
// Replaces the per-page `tifffile.TiffFile(...).pages[i].asarray()` of the reference's lazy tile
// loader (src/magnify/reader.py:265-279, one dask chunk = one TIFF page) on the staging side of
// the hot path.  Host code only: the IFD chain is parsed once per
// file, strips are coalesced into as few pread(2) calls as possible and land directly in the
// destination (a pinned staging buffer), pages are spread over a small thread pool.  Layout
// knowledge follows the TIFF 6.0 specification and the BigTIFF extension (magic 43, 8-byte
// offsets).  Uncompressed stripped pages -- what Micro-Manager and most acquisition software write
// -- take the direct path; LZW, Deflate and PackBits data, horizontal differencing (Predictor 2)
// and tiled pages are decoded chunk by chunk on the reading thread.  Anything else (JPEG, float
// predictor, planar multi-sample) is reported, never silently mis-read.
#include "magnify_b200.h"

#include <fcntl.h>
#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Page {
  int64_t width = 0, height = 0;
  int bits = 1, samples = 1, sample_format = 1, compression = 1, planar = 1, predictor = 1;
  int photometric = 1;
  int64_t rows_per_strip = -1;
  bool tiled = false;
  int64_t tile_width = 0, tile_length = 0;
  int64_t subfile_type = 0;
  std::vector<uint64_t> offsets, counts;
  uint64_t desc_offset = 0, desc_len = 0;   // ImageDescription (tag 270) bytes in the file
  int64_t row_bytes() const { return (width * samples * bits + 7) / 8; }
  int64_t data_bytes() const { return row_bytes() * height; }
};

struct TiffFile {
  int fd = -1;
  bool big = false, swap = false;   // BigTIFF; file byte order != host byte order
  uint64_t file_size = 0;
  std::vector<Page> pages;
  int error = 0;
};

bool host_is_little() {
  const uint16_t one = 1;
  return *reinterpret_cast<const uint8_t*>(&one) == 1;
}

bool pread_all(int fd, void* dst, uint64_t n, uint64_t off) {
  uint8_t* p = static_cast<uint8_t*>(dst);
  while (n > 0) {
    ssize_t got = ::pread(fd, p, n, static_cast<off_t>(off));
    if (got < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (got == 0) return false;   // short file
    p += got;
    off += static_cast<uint64_t>(got);
    n -= static_cast<uint64_t>(got);
  }
  return true;
}

template <typename T>
T load(const uint8_t* p, bool swap) {
  T v;
  std::memcpy(&v, p, sizeof(T));
  if (swap) {
    uint8_t* b = reinterpret_cast<uint8_t*>(&v);
    std::reverse(b, b + sizeof(T));
  }
  return v;
}

int type_size(int type) {
  switch (type) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    default: return 0;
  }
}

// Integer value `i` of an entry's value array (types BYTE/SHORT/LONG/LONG8 and signed kin).
uint64_t value_at(const uint8_t* p, int type, uint64_t i, bool swap) {
  switch (type) {
    case 1: case 6: case 7: return p[i];
    case 3: case 8: return load<uint16_t>(p + 2 * i, swap);
    case 4: case 9: case 13: return load<uint32_t>(p + 4 * i, swap);
    case 16: case 17: case 18: return load<uint64_t>(p + 8 * i, swap);
    default: return 0;
  }
}

// Parse the IFD at `off`; returns the offset of the next IFD (0 = end) or UINT64_MAX on error.
uint64_t parse_ifd(TiffFile& f, uint64_t off, Page& page) {
  const uint64_t bad = UINT64_MAX;
  const int count_bytes = f.big ? 8 : 2, entry_bytes = f.big ? 20 : 12, next_bytes = f.big ? 8 : 4;
  uint8_t head[8];
  if (off + count_bytes > f.file_size || !pread_all(f.fd, head, count_bytes, off)) return bad;
  const uint64_t n = f.big ? load<uint64_t>(head, f.swap) : load<uint16_t>(head, f.swap);
  if (n == 0 || n > 65536) return bad;
  std::vector<uint8_t> buf(n * entry_bytes + next_bytes);
  if (off + count_bytes + buf.size() > f.file_size ||
      !pread_all(f.fd, buf.data(), buf.size(), off + count_bytes))
    return bad;
  std::vector<uint8_t> ext;
  for (uint64_t e = 0; e < n; ++e) {
    const uint8_t* p = buf.data() + e * entry_bytes;
    const int tag = load<uint16_t>(p, f.swap), type = load<uint16_t>(p + 2, f.swap);
    const uint64_t count = f.big ? load<uint64_t>(p + 4, f.swap) : load<uint32_t>(p + 4, f.swap);
    const uint8_t* inl = p + (f.big ? 12 : 8);
    const uint64_t inline_cap = f.big ? 8 : 4;
    const int ts = type_size(type);
    if (ts == 0) continue;   // unknown field type: skip the field (TIFF 6.0 p.16)
    if (count > (UINT64_MAX >> 4)) return bad;
    const uint64_t bytes = count * ts;
    const bool inl_value = bytes <= inline_cap;
    const uint64_t where = inl_value ? 0 : (f.big ? load<uint64_t>(inl, f.swap) : load<uint32_t>(inl, f.swap));
    auto fetch = [&]() -> const uint8_t* {
      if (inl_value) return inl;
      if (where + bytes > f.file_size) return nullptr;
      ext.resize(bytes);
      return pread_all(f.fd, ext.data(), bytes, where) ? ext.data() : nullptr;
    };
    auto scalar = [&](uint64_t& out) -> bool {
      const uint8_t* v = fetch();
      if (!v || count < 1) return false;
      out = value_at(v, type, 0, f.swap);
      return true;
    };
    uint64_t v = 0;
    switch (tag) {
      case 254: if (!scalar(v)) return bad; page.subfile_type = static_cast<int64_t>(v); break;
      case 256: if (!scalar(v)) return bad; page.width = static_cast<int64_t>(v); break;
      case 257: if (!scalar(v)) return bad; page.height = static_cast<int64_t>(v); break;
      case 258: {   // BitsPerSample: one value per sample, all equal for the pages handled here
        const uint8_t* a = fetch();
        if (!a || count < 1) return bad;
        page.bits = static_cast<int>(value_at(a, type, 0, f.swap));
        for (uint64_t i = 1; i < count; ++i)
          if (static_cast<int>(value_at(a, type, i, f.swap)) != page.bits) page.compression = -1;
        break;
      }
      case 259: if (!scalar(v)) return bad; if (page.compression != -1) page.compression = static_cast<int>(v); break;
      case 262: if (!scalar(v)) return bad; page.photometric = static_cast<int>(v); break;
      case 270:
        page.desc_len = bytes;
        if (inl_value) {   // a description of <= 4 (8) bytes lives inside the entry itself
          page.desc_offset = off + count_bytes + e * entry_bytes + (f.big ? 12 : 8);
        } else if (where + bytes <= f.file_size) {
          page.desc_offset = where;
        } else {
          page.desc_len = 0;   // points past the end of the file: treat as absent
        }
        break;
      case 273: case 324: case 279: case 325: {
        const uint8_t* a = fetch();
        if (!a) return bad;
        std::vector<uint64_t>& dst = (tag == 273 || tag == 324) ? page.offsets : page.counts;
        dst.resize(count);
        for (uint64_t i = 0; i < count; ++i) dst[i] = value_at(a, type, i, f.swap);
        if (tag == 324 || tag == 325) page.tiled = true;
        break;
      }
      case 277: if (!scalar(v)) return bad; page.samples = static_cast<int>(v); break;
      case 278: if (!scalar(v)) return bad; page.rows_per_strip = static_cast<int64_t>(v); break;
      case 284: if (!scalar(v)) return bad; page.planar = static_cast<int>(v); break;
      case 317: if (!scalar(v)) return bad; page.predictor = static_cast<int>(v); break;
      case 322: if (!scalar(v)) return bad; page.tiled = true; page.tile_width = static_cast<int64_t>(v); break;
      case 323: if (!scalar(v)) return bad; page.tiled = true; page.tile_length = static_cast<int64_t>(v); break;
      case 339: if (!scalar(v)) return bad; page.sample_format = static_cast<int>(v); break;
      default: break;
    }
  }
  const uint8_t* nx = buf.data() + n * entry_bytes;
  return f.big ? load<uint64_t>(nx, f.swap) : load<uint32_t>(nx, f.swap);
}

int open_file(const char* path, TiffFile& f, int64_t max_pages) {
  f.fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (f.fd < 0) return MGB_EIO;
  struct stat st;
  if (::fstat(f.fd, &st) != 0) return MGB_EIO;
  f.file_size = static_cast<uint64_t>(st.st_size);
  uint8_t h[16];
  if (f.file_size < 8 || !pread_all(f.fd, h, 8, 0)) return MGB_EFORMAT;
  bool little;
  if (h[0] == 'I' && h[1] == 'I') little = true;
  else if (h[0] == 'M' && h[1] == 'M') little = false;
  else return MGB_EFORMAT;
  f.swap = little != host_is_little();
  const uint16_t magic = load<uint16_t>(h + 2, f.swap);
  uint64_t off;
  if (magic == 42) {
    f.big = false;
    off = load<uint32_t>(h + 4, f.swap);
  } else if (magic == 43) {
    f.big = true;
    if (f.file_size < 16 || !pread_all(f.fd, h, 16, 0)) return MGB_EFORMAT;
    if (load<uint16_t>(h + 4, f.swap) != 8 || load<uint16_t>(h + 6, f.swap) != 0) return MGB_EFORMAT;
    off = load<uint64_t>(h + 8, f.swap);
  } else {
    return MGB_EFORMAT;
  }
  // walk the main IFD chain (one page per IFD, like tifffile's `pages`); a cycle ends the walk
  std::vector<uint64_t> seen;
  while (off != 0 && (max_pages < 0 || static_cast<int64_t>(f.pages.size()) < max_pages)) {
    if (std::find(seen.begin(), seen.end(), off) != seen.end()) break;
    seen.push_back(off);
    Page page;
    const uint64_t next = parse_ifd(f, off, page);
    if (next == UINT64_MAX) return MGB_EFORMAT;
    if (page.rows_per_strip <= 0 || page.rows_per_strip > page.height) page.rows_per_strip = page.height;
    f.pages.push_back(std::move(page));
    off = next;
  }
  return f.pages.empty() ? MGB_EFORMAT : MGB_OK;
}

bool plain_strips(const Page& p) { return p.compression == 1 && p.predictor == 1 && !p.tiled; }

// Why a page cannot be decoded by this reader (MGB_OK when it can).
int page_supported(const Page& p) {
  const bool codec = p.compression == 1 || p.compression == 5 || p.compression == 8 || p.compression == 32946 ||
                     p.compression == 32773;
  if (!codec) return MGB_EUNSUPPORTED;
  if (p.predictor != 1 && !(p.predictor == 2 && p.sample_format != 3)) return MGB_EUNSUPPORTED;
  if (p.width <= 0 || p.height <= 0) return MGB_EUNSUPPORTED;
  if (p.bits != 8 && p.bits != 16 && p.bits != 32 && p.bits != 64) return MGB_EUNSUPPORTED;
  if (p.samples != 1 && p.planar != 1) return MGB_EUNSUPPORTED;
  if (p.offsets.empty() || p.offsets.size() != p.counts.size()) return MGB_EFORMAT;
  if (p.tiled) {
    if (p.tile_width <= 0 || p.tile_length <= 0 || p.tile_width > (1 << 20) || p.tile_length > (1 << 20)) return MGB_EFORMAT;
    const int64_t across = (p.width + p.tile_width - 1) / p.tile_width, down = (p.height + p.tile_length - 1) / p.tile_length;
    if (static_cast<int64_t>(p.offsets.size()) != across * down) return MGB_EFORMAT;
  } else {
    const int64_t strips = (p.height + p.rows_per_strip - 1) / p.rows_per_strip;
    if (static_cast<int64_t>(p.offsets.size()) != strips) return MGB_EFORMAT;
  }
  return MGB_OK;
}

void swap_inplace(uint8_t* p, int64_t bytes, int itemsize) {
  if (itemsize == 2) {
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    for (int64_t i = 0; i < bytes / 2; ++i) q[i] = static_cast<uint16_t>((q[i] >> 8) | (q[i] << 8));
  } else if (itemsize == 4) {
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    for (int64_t i = 0; i < bytes / 4; ++i) q[i] = __builtin_bswap32(q[i]);
  } else if (itemsize == 8) {
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
    for (int64_t i = 0; i < bytes / 8; ++i) q[i] = __builtin_bswap64(q[i]);
  }
}

// ---- decoders: each fills exactly `want` bytes of dst (false: corrupt or short data) -----------

// TIFF 6.0 section 13: variable-width codes packed most significant bit first, ClearCode 256,
// EndOfInformation 257, first free code 258, the width grows one code early (at 511, 1023, 2047).
// Every table string is a substring of what has already been written: string(next) =
// string(prev) + first byte of the current string, and those bytes sit back to back in the output.
// So an entry is just (position, length) into the output and emitting a code is a short forward
// copy (byte-wise, because the "KwKwK" code may overlap its own source).
struct LzwTable {
  uint32_t pos[4096];
  uint16_t len[4096];
};

bool lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
  static thread_local LzwTable table;
  LzwTable& t = table;
  size_t out = 0, in = 0;
  uint64_t acc = 0;       // bit reservoir, most significant bits first
  int have = 0;
  int width = 9, next = 258;
  bool has_prev = false;
  size_t prev_pos = 0, prev_len = 0;
  while (out < want) {
    while (have <= 56 && in < n) {
      acc |= static_cast<uint64_t>(src[in++]) << (56 - have);
      have += 8;
    }
    if (have < width) return false;
    const int code = static_cast<int>(acc >> (64 - width));
    acc <<= width;
    have -= width;
    if (code == 257) break;
    if (code == 256) {
      width = 9;
      next = 258;
      has_prev = false;
      continue;
    }
    size_t from, len;
    const size_t here = out;
    if (code < 256) {
      dst[out++] = static_cast<uint8_t>(code);
      len = 1;
    } else {
      if (!has_prev) return false;
      if (code < next) {
        if (code < 258) return false;
        from = t.pos[code];
        len = t.len[code];
      } else if (code == next) {
        from = prev_pos;
        len = prev_len + 1;
      } else {
        return false;
      }
      const size_t take = std::min(len, want - out);
      for (size_t k = 0; k < take; ++k) dst[out + k] = dst[from + k];
      out += take;
    }
    if (has_prev && next < 4096) {
      t.pos[next] = static_cast<uint32_t>(prev_pos);
      t.len[next] = static_cast<uint16_t>(prev_len + 1);
      ++next;
    }
    has_prev = true;
    prev_pos = here;
    prev_len = len;
    if (next == (1 << width) - 1 && width < 12) ++width;
  }
  return out == want;
}

bool packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
  size_t in = 0, out = 0;
  while (out < want && in < n) {
    const int8_t h = static_cast<int8_t>(src[in++]);
    if (h >= 0) {
      const size_t run = static_cast<size_t>(h) + 1;
      if (in + run > n) return false;
      const size_t take = std::min(run, want - out);
      std::memcpy(dst + out, src + in, take);
      in += run;
      out += take;
    } else if (h != -128) {
      if (in >= n) return false;
      const size_t run = static_cast<size_t>(1 - h), take = std::min(run, want - out);
      std::memset(dst + out, src[in++], take);
      out += take;
    }
  }
  return out == want;
}

bool inflate_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t want) {
  z_stream z;
  std::memset(&z, 0, sizeof(z));
  if (inflateInit(&z) != Z_OK) return false;
  z.next_in = const_cast<Bytef*>(src);
  z.avail_in = static_cast<uInt>(n);
  z.next_out = dst;
  z.avail_out = static_cast<uInt>(want);
  const int rc = inflate(&z, Z_FINISH);
  const bool ok = (rc == Z_STREAM_END || rc == Z_OK || rc == Z_BUF_ERROR) && z.avail_out == 0;
  inflateEnd(&z);
  return ok;
}

// Predictor 2 (TIFF 6.0 section 14): every sample is stored as the difference to the sample
// `samples` positions earlier in its row; undo it in native byte order.
template <typename T>
void undo_differencing(uint8_t* row, int64_t items, int samples) {
  T* q = reinterpret_cast<T*>(row);
  for (int64_t i = samples; i < items; ++i) q[i] = static_cast<T>(q[i] + q[i - samples]);
}

// Decode one strip or tile of `rows` x `cols` pixels into `chunk` (rows * cols * samples items,
// native byte order, differencing undone).
int decode_chunk(const TiffFile& f, const Page& p, uint64_t offset, uint64_t count, int64_t rows, int64_t cols,
                 std::vector<uint8_t>& packed, uint8_t* chunk) {
  const int item = p.bits / 8;
  const size_t want = static_cast<size_t>(rows * cols * p.samples * item);
  if (offset + count > f.file_size || count > (1ull << 31) || want > (1ull << 31)) return MGB_EFORMAT;
  if (p.compression == 1) {
    if (count < want) return MGB_EFORMAT;
    if (!pread_all(f.fd, chunk, want, offset)) return MGB_EIO;
  } else {
    packed.resize(count);
    if (!pread_all(f.fd, packed.data(), count, offset)) return MGB_EIO;
    bool ok;
    if (p.compression == 5) ok = lzw_decode(packed.data(), count, chunk, want);
    else if (p.compression == 32773) ok = packbits_decode(packed.data(), count, chunk, want);
    else ok = inflate_decode(packed.data(), count, chunk, want);
    if (!ok) return MGB_EFORMAT;
  }
  if (f.swap && item > 1) swap_inplace(chunk, static_cast<int64_t>(want), item);
  if (p.predictor == 2) {
    const int64_t items = cols * p.samples;
    for (int64_t r = 0; r < rows; ++r) {
      uint8_t* row = chunk + r * items * item;
      if (item == 1) undo_differencing<uint8_t>(row, items, p.samples);
      else if (item == 2) undo_differencing<uint16_t>(row, items, p.samples);
      else if (item == 4) undo_differencing<uint32_t>(row, items, p.samples);
      else undo_differencing<uint64_t>(row, items, p.samples);
    }
  }
  return MGB_OK;
}

int read_page(const TiffFile& f, const Page& p, void* dst, int64_t dst_bytes) {
  const int ok = page_supported(p);
  if (ok != MGB_OK) return ok;
  const int64_t row_bytes = p.row_bytes(), total = p.data_bytes();
  if (dst_bytes < total) return MGB_EINVAL;
  uint8_t* out = static_cast<uint8_t*>(dst);
  if (!plain_strips(p)) {
    // compressed, differenced or tiled pages: chunk by chunk through a scratch buffer
    std::vector<uint8_t> packed, chunk;
    const int item = p.bits / 8;
    if (p.tiled) {
      const int64_t across = (p.width + p.tile_width - 1) / p.tile_width;
      const int64_t tile_row_bytes = p.tile_width * p.samples * item;
      chunk.resize(static_cast<size_t>(p.tile_length * tile_row_bytes));
      for (size_t t = 0; t < p.offsets.size(); ++t) {
        const int rc = decode_chunk(f, p, p.offsets[t], p.counts[t], p.tile_length, p.tile_width, packed, chunk.data());
        if (rc != MGB_OK) return rc;
        const int64_t y0 = static_cast<int64_t>(t / across) * p.tile_length, x0 = static_cast<int64_t>(t % across) * p.tile_width;
        const int64_t rows = std::min(p.tile_length, p.height - y0), cols = std::min(p.tile_width, p.width - x0);
        for (int64_t r = 0; r < rows; ++r)
          std::memcpy(out + (y0 + r) * row_bytes + x0 * p.samples * item, chunk.data() + r * tile_row_bytes,
                      static_cast<size_t>(cols * p.samples * item));
      }
    } else {
      for (size_t s = 0; s < p.offsets.size(); ++s) {
        const int64_t y0 = static_cast<int64_t>(s) * p.rows_per_strip;
        const int64_t rows = std::min<int64_t>(p.rows_per_strip, p.height - y0);
        const int rc = decode_chunk(f, p, p.offsets[s], p.counts[s], rows, p.width, packed, out + y0 * row_bytes);
        if (rc != MGB_OK) return rc;
      }
    }
    return MGB_OK;
  }
  const int64_t strips = static_cast<int64_t>(p.offsets.size());
  int64_t s = 0;
  while (s < strips) {
    // coalesce strips that are back to back in the file and full-sized into one pread
    const int64_t first = s;
    const uint64_t start = p.offsets[s];
    uint64_t run = 0;
    while (s < strips) {
      const int64_t rows = std::min<int64_t>(p.rows_per_strip, p.height - s * p.rows_per_strip);
      const uint64_t want = static_cast<uint64_t>(rows * row_bytes);
      if (p.counts[s] < want) return MGB_EFORMAT;   // truncated strip
      if (p.offsets[s] != start + run) break;
      run += want;
      ++s;
      if (p.counts[s - 1] != want) break;   // padded strip: the next one cannot be contiguous data
    }
    if (start + run > f.file_size) return MGB_EFORMAT;
    if (!pread_all(f.fd, out + first * p.rows_per_strip * row_bytes, run, start)) return MGB_EIO;
  }
  if (f.swap && p.bits > 8) swap_inplace(out, total, p.bits / 8);
  return MGB_OK;
}

template <typename Fn>
int parallel_for(int64_t n, int threads, Fn fn) {
  std::atomic<int64_t> next{0};
  std::atomic<int> status{MGB_OK};
  auto worker = [&]() {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n || status.load() != MGB_OK) return;
      int rc;
      try {
        rc = fn(i);
      } catch (...) {   // e.g. bad_alloc on a corrupt directory: never let it cross the C ABI / thread
        rc = MGB_EIO;
      }
      if (rc != MGB_OK) {
        int expected = MGB_OK;
        status.compare_exchange_strong(expected, rc);
      }
    }
  };
  const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, n)));
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return status.load();
}

}  // namespace

extern "C" {

int mgb_tiff_open(const char* host_path, void** host_handle) {
  if (!host_path || !host_handle) return MGB_EINVAL;
  TiffFile* f = new (std::nothrow) TiffFile();
  if (!f) return MGB_EIO;
  int rc;
  try {
    rc = open_file(host_path, *f, -1);
  } catch (...) {
    rc = MGB_EIO;
  }
  if (rc != MGB_OK) {
    if (f->fd >= 0) ::close(f->fd);
    delete f;
    *host_handle = nullptr;
    return rc;
  }
  *host_handle = f;
  return MGB_OK;
}

int mgb_tiff_close(void* host_handle) {
  TiffFile* f = static_cast<TiffFile*>(host_handle);
  if (!f) return MGB_EINVAL;
  if (f->fd >= 0) ::close(f->fd);
  delete f;
  return MGB_OK;
}

int mgb_tiff_page_count(const void* host_handle, int64_t* host_count) {
  const TiffFile* f = static_cast<const TiffFile*>(host_handle);
  if (!f || !host_count) return MGB_EINVAL;
  *host_count = static_cast<int64_t>(f->pages.size());
  return MGB_OK;
}

int mgb_tiff_page_info(const void* host_handle, int64_t page, int64_t* host_info) {
  const TiffFile* f = static_cast<const TiffFile*>(host_handle);
  if (!f || !host_info || page < 0 || page >= static_cast<int64_t>(f->pages.size())) return MGB_EINVAL;
  const Page& p = f->pages[page];
  host_info[0] = p.width;
  host_info[1] = p.height;
  host_info[2] = p.bits;
  host_info[3] = p.samples;
  host_info[4] = p.sample_format;
  host_info[5] = p.compression;
  host_info[6] = p.data_bytes();
  host_info[7] = page_supported(p);
  host_info[8] = static_cast<int64_t>(p.desc_len);
  host_info[9] = static_cast<int64_t>(p.offsets.size());
  host_info[10] = f->big ? 1 : 0;
  host_info[11] = f->swap == host_is_little() ? 1 : 0;   // 1 = big-endian file
  return MGB_OK;
}

int mgb_tiff_description(const void* host_handle, int64_t page, char* host_buf, int64_t capacity) {
  const TiffFile* f = static_cast<const TiffFile*>(host_handle);
  if (!f || !host_buf || page < 0 || page >= static_cast<int64_t>(f->pages.size())) return MGB_EINVAL;
  const Page& p = f->pages[page];
  if (capacity < static_cast<int64_t>(p.desc_len)) return MGB_EINVAL;
  if (p.desc_len == 0) return MGB_OK;
  if (p.desc_offset + p.desc_len > f->file_size) return MGB_EFORMAT;
  return pread_all(f->fd, host_buf, p.desc_len, p.desc_offset) ? MGB_OK : MGB_EIO;
}

int mgb_tiff_read_pages(const void* host_handle, const int64_t* host_pages, int64_t n_pages, void* host_dst,
                        int64_t dst_stride_bytes, int threads) {
  const TiffFile* f = static_cast<const TiffFile*>(host_handle);
  if (!f || !host_pages || !host_dst || n_pages < 0 || dst_stride_bytes < 0) return MGB_EINVAL;
  for (int64_t i = 0; i < n_pages; ++i) {
    if (host_pages[i] < 0 || host_pages[i] >= static_cast<int64_t>(f->pages.size())) return MGB_EINVAL;
    if (f->pages[host_pages[i]].data_bytes() > dst_stride_bytes) return MGB_EINVAL;
  }
  uint8_t* dst = static_cast<uint8_t*>(host_dst);
  return parallel_for(n_pages, threads, [&](int64_t i) {
    return read_page(*f, f->pages[host_pages[i]], dst + i * dst_stride_bytes, dst_stride_bytes);
  });
}

int mgb_tiff_read_files(const char* const* host_paths, int64_t n_files, int64_t page, int64_t width, int64_t height,
                        int bits, void* host_dst, int64_t dst_stride_bytes, int threads) {
  if (!host_paths || !host_dst || n_files < 0 || page < 0 || dst_stride_bytes < 0) return MGB_EINVAL;
  uint8_t* dst = static_cast<uint8_t*>(host_dst);
  return parallel_for(n_files, threads, [&](int64_t i) {
    TiffFile f;
    int rc = host_paths[i] ? open_file(host_paths[i], f, page + 1) : MGB_EINVAL;
    if (rc == MGB_OK) {
      if (page >= static_cast<int64_t>(f.pages.size())) {
        rc = MGB_EINVAL;
      } else {
        const Page& p = f.pages[page];
        // every file of one acquisition must hold pages of the announced geometry
        if (p.width != width || p.height != height || p.bits != bits || p.samples != 1) rc = MGB_EFORMAT;
        else rc = read_page(f, p, dst + i * dst_stride_bytes, dst_stride_bytes);
      }
    }
    if (f.fd >= 0) ::close(f.fd);
    return rc;
  });
}

}  // extern "C"

// ---- writer: uncompressed little-endian pages, one strip each, directories after the pixel data ----
namespace {

bool pwrite_all(int fd, const void* src, uint64_t n, uint64_t off) {
  const uint8_t* p = static_cast<const uint8_t*>(src);
  while (n > 0) {
    const ssize_t put = ::pwrite(fd, p, n, static_cast<off_t>(off));
    if (put < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += put;
    off += static_cast<uint64_t>(put);
    n -= static_cast<uint64_t>(put);
  }
  return true;
}

template <typename T>
void put(std::vector<uint8_t>& buf, T v) {   // host is little-endian on every platform this library builds for
  const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
  buf.insert(buf.end(), b, b + sizeof(T));
}

void put_entry(std::vector<uint8_t>& buf, bool big, uint16_t tag, uint16_t type, uint64_t count, uint64_t value) {
  put<uint16_t>(buf, tag);
  put<uint16_t>(buf, type);
  if (big) {
    put<uint64_t>(buf, count);
    put<uint64_t>(buf, value);
  } else {
    put<uint32_t>(buf, static_cast<uint32_t>(count));
    put<uint32_t>(buf, static_cast<uint32_t>(value));
  }
}

}  // namespace

extern "C" int mgb_tiff_write(const char* host_path, const void* host_pages, int64_t n_pages, int64_t height,
                              int64_t width, int bits, int sample_format, int bigtiff, const char* host_description,
                              int threads) {
  if (!host_path || !host_pages || n_pages <= 0 || height <= 0 || width <= 0) return MGB_EINVAL;
  if ((bits != 8 && bits != 16 && bits != 32 && bits != 64) || sample_format < 1 || sample_format > 3) return MGB_EINVAL;
  if (!host_is_little()) return MGB_EUNSUPPORTED;
  const uint64_t page_bytes = static_cast<uint64_t>(height) * width * (bits / 8);
  const uint64_t desc_len = host_description ? std::strlen(host_description) + 1 : 0;
  const uint64_t header = 16;                                   // classic files simply leave 8 bytes unused
  const uint64_t desc_off = header + page_bytes * n_pages;      // description, then the directories
  const bool big = bigtiff > 0 || (bigtiff < 0 && desc_off + desc_len + 256ull * n_pages > 0xfff00000ull);
  if (!big && desc_off + desc_len + 256ull * n_pages > 0xfff00000ull) return MGB_EINVAL;   // needs BigTIFF
  const int n_entries = host_description ? 11 : 10;
  const uint64_t ifd_bytes = (big ? 8 : 2) + static_cast<uint64_t>(n_entries) * (big ? 20 : 12) + (big ? 8 : 4);
  uint64_t ifd0 = desc_off + desc_len;
  ifd0 += ifd0 & 1;                                             // directories start on an even offset
  int rc = MGB_OK;
  const int fd = ::open(host_path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
  if (fd < 0) return MGB_EIO;
  try {
    std::vector<uint8_t> head;
    head.push_back('I');
    head.push_back('I');
    if (big) {
      put<uint16_t>(head, 43);
      put<uint16_t>(head, 8);
      put<uint16_t>(head, 0);
      put<uint64_t>(head, ifd0);
    } else {
      put<uint16_t>(head, 42);
      put<uint32_t>(head, static_cast<uint32_t>(ifd0));
      head.resize(16, 0);
    }
    std::vector<uint8_t> dirs;
    for (int64_t k = 0; k < n_pages; ++k) {
      if (big) put<uint64_t>(dirs, static_cast<uint64_t>(n_entries));
      else put<uint16_t>(dirs, static_cast<uint16_t>(n_entries));
      const bool with_desc = host_description && k == 0;
      put_entry(dirs, big, 256, 4, 1, static_cast<uint64_t>(width));
      put_entry(dirs, big, 257, 4, 1, static_cast<uint64_t>(height));
      put_entry(dirs, big, 258, 3, 1, static_cast<uint64_t>(bits));
      put_entry(dirs, big, 259, 3, 1, 1);
      put_entry(dirs, big, 262, 3, 1, 1);
      if (host_description) {   // later pages carry a one-byte (empty) description so that every directory has the same size
        if (with_desc && desc_len > (big ? 8u : 4u)) {
          put_entry(dirs, big, 270, 2, desc_len, desc_off);
        } else if (with_desc) {   // short enough to live in the value field itself
          uint64_t packed = 0;
          std::memcpy(&packed, host_description, desc_len);
          put_entry(dirs, big, 270, 2, desc_len, packed);
        } else {
          put_entry(dirs, big, 270, 2, 1, 0);
        }
      }
      put_entry(dirs, big, 273, big ? 16 : 4, 1, header + page_bytes * k);
      put_entry(dirs, big, 277, 3, 1, 1);
      put_entry(dirs, big, 278, 4, 1, static_cast<uint64_t>(height));
      put_entry(dirs, big, 279, big ? 16 : 4, 1, page_bytes);
      put_entry(dirs, big, 339, 3, 1, static_cast<uint64_t>(sample_format));
      const uint64_t next = k + 1 < n_pages ? ifd0 + ifd_bytes * (k + 1) : 0;
      if (big) put<uint64_t>(dirs, next);
      else put<uint32_t>(dirs, static_cast<uint32_t>(next));
    }
    if (!pwrite_all(fd, head.data(), head.size(), 0)) rc = MGB_EIO;
    if (rc == MGB_OK && desc_len && !pwrite_all(fd, host_description, desc_len, desc_off)) rc = MGB_EIO;
    if (rc == MGB_OK && !pwrite_all(fd, dirs.data(), dirs.size(), ifd0)) rc = MGB_EIO;
    if (rc == MGB_OK) {
      const uint8_t* src = static_cast<const uint8_t*>(host_pages);
      rc = parallel_for(n_pages, threads, [&](int64_t k) {
        return pwrite_all(fd, src + page_bytes * k, page_bytes, header + page_bytes * k) ? MGB_OK : MGB_EIO;
      });
    }
  } catch (...) {
    rc = MGB_EIO;
  }
  if (::close(fd) != 0 && rc == MGB_OK) rc = MGB_EIO;
  return rc;
}
