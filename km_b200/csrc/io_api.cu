// C ABI of libkm_b200.so, part 2: the files either side of the table -- FASTA / FASTQ reads in
// (km_table_count_file, the input side of `jellyfish count`, example/run_leucegene.sh:22), Jellyfish
// binary/sorted databases in and out (km_table_open_jf: km/utils/Jellyfish.py:23-45; km_table_write_jf).
#include "host_common.h"
#include "table.h"
#include "exec_model.h"
#include <sys/types.h>

// ---- counting straight from FASTA / FASTQ files (plain or .gz) ------------------------------------------
// zlib is looked up at run time (dlopen), like the driver's virtual-memory entry points: the library loads on
// a machine without it and only .gz input is refused there.
#include <dlfcn.h>
struct ZLib {
    void* (*open)(const char*, const char*) = nullptr;
    int (*read)(void*, void*, unsigned) = nullptr;
    int (*close)(void*) = nullptr;
    int (*buffer)(void*, unsigned) = nullptr;
    bool tried = false, ok = false;
};
static ZLib g_z;
static bool zlib_load() {
    if (g_z.tried) return g_z.ok;
    g_z.tried = true;
    void* h = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
    g_z.open = (void* (*)(const char*, const char*))dlsym(h, "gzopen");
    g_z.read = (int (*)(void*, void*, unsigned))dlsym(h, "gzread");
    g_z.close = (int (*)(void*))dlsym(h, "gzclose");
    g_z.buffer = (int (*)(void*, unsigned))dlsym(h, "gzbuffer");
    g_z.ok = g_z.open && g_z.read && g_z.close;
    return g_z.ok;
}
struct LineReader {                 // lines of a plain or gzip file, without their line ends
    FILE* f = nullptr; void* gz = nullptr;
    std::vector<char> buf; size_t pos = 0, end = 0; bool eof = false;
    bool fill() {
        if (eof) return false;
        if (pos > 0) { memmove(buf.data(), buf.data() + pos, end - pos); end -= pos; pos = 0; }
        if (end == buf.size()) buf.resize(buf.size() * 2);
        const size_t room = buf.size() - end;
        long got = gz ? (long)g_z.read(gz, buf.data() + end, (unsigned)std::min<size_t>(room, 1u << 30)) : (long)fread(buf.data() + end, 1, room, f);
        if (got <= 0) { eof = true; return false; }
        end += (size_t)got;
        return true;
    }
    // next line into (*p, *n); false at end of file
    bool next(const char** p, size_t* n) {
        for (;;) {
            const char* nl = (const char*)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                *p = buf.data() + pos; *n = (size_t)(nl - *p);
                pos = (size_t)(nl - buf.data()) + 1;
                if (*n && (*p)[*n - 1] == '\r') --*n;
                return true;
            }
            if (!fill()) {
                if (pos < end) { *p = buf.data() + pos; *n = end - pos; pos = end; if (*n && (*p)[*n - 1] == '\r') --*n; return true; }
                return false;
            }
        }
    }
};

// `jellyfish count` input side.  The file's sequence lines are copied ONCE, straight into a pinned staging buffer
// (a newline between sequences), FASTQ quality lines into a parallel buffer at the same offsets; the -Q mask, the
// 2-bit packing and the counting all happen on the device (km_count_text_kernel).  Two staging buffers: while the
// GPU copies and counts one, this thread parses the next (CountStream).  FASTQ records are four lines; FASTA
// sequences may span lines.
extern "C" int km_table_count_file(km_table* t, const char* path, int min_qual_char, uint64_t* n_reads_out, uint64_t* n_bases_out) {
    if (!t || !path) return fail(KM_E_ARG, "km_table_count_file: bad argument");
    if (t->lines) return fail(KM_E_ARG, "km_table_count_file: counting needs the sector-bucket layout");
    LineReader R;
    const size_t plen = strlen(path);
    const bool gz = plen > 3 && strcmp(path + plen - 3, ".gz") == 0;
    if (gz) {
        if (!zlib_load()) return fail(KM_E_IO, "%s: libz.so.1 not found, decompress the file first", path);
        R.gz = g_z.open(path, "rb");
        if (!R.gz) return fail(KM_E_IO, "cannot open %s", path);
        if (g_z.buffer) g_z.buffer(R.gz, 1u << 20);
    } else {
        R.f = strcmp(path, "-") == 0 ? stdin : fopen(path, "rb");
        if (!R.f) return fail(KM_E_IO, "cannot open %s", path);
    }
    R.buf.resize((size_t)8 << 20);
    CountStream cs;
    const bool want_q = min_qual_char > 0;
    int rc = cs.open(t, (size_t)32 << 20, want_q);
    size_t fill = 0;
    const size_t k1 = (size_t)t->k - 1;
    uint64_t n_reads = 0, n_bases = 0;
    // append `n` bases (and their qualities) to the stream; a piece that does not fit goes out and the sequence
    // continues k - 1 bases back in the next buffer, so no k-mer is lost or counted twice
    auto put = [&](const char* sq, const char* ql, size_t n) -> int {
        while (n) {
            const size_t room = fill + 1 < cs.cap ? cs.cap - fill - 1 : 0;
            if (n <= room) {
                memcpy(cs.seq() + fill, sq, n);
                if (want_q) { if (ql) memcpy(cs.qual() + fill, ql, n); else memset(cs.qual() + fill, 0x7F, n); }
                fill += n; n = 0;
            } else if (fill == 0) {
                memcpy(cs.seq(), sq, room);
                if (want_q) { if (ql) memcpy(cs.qual(), ql, room); else memset(cs.qual(), 0x7F, room); }
                if (int r = cs.submit(room, min_qual_char)) return r;
                sq += room - k1; if (ql) ql += room - k1; n -= room - k1;
            } else { if (int r = cs.submit(fill, min_qual_char)) return r; fill = 0; }
        }
        return 0;
    };
    auto end_seq = [&]() { cs.seq()[fill] = '\n'; if (want_q) cs.qual()[fill] = 0x7F; ++fill; };
    const char* ln; size_t n;
    bool first = true, fastq = false, open_seq = false;
    std::vector<char> held;                       // FASTQ: the sequence line, kept while its quality line is fetched
    while (!rc && R.next(&ln, &n)) {
        if (first) {
            if (!n) continue;
            first = false;
            if (ln[0] == '@') fastq = true;
            else if (ln[0] != '>') { rc = fail(KM_E_IO, "%s: neither FASTA nor FASTQ", path); break; }
        }
        if (fastq) {
            if (!n) continue;                              // stray blank line between records
            // ln is the header; then sequence, '+', quality
            const char* sq; size_t sn;
            if (!R.next(&sq, &sn)) break;
            held.assign(sq, sq + sn);                      // (a line stays valid only until the next call of next())
            const char* pl; size_t pn; const char* ql = nullptr; size_t qn = 0;
            const bool whole = R.next(&pl, &pn) && R.next(&ql, &qn);
            if (sn) {
                rc = put(held.data(), whole && want_q && qn == sn ? ql : nullptr, sn);
                if (!rc) { end_seq(); n_reads += 1; n_bases += sn; }
            }
            if (!whole) break;
        } else {
            if (n && ln[0] == '>') { if (open_seq) { end_seq(); open_seq = false; } }
            else if (n) {
                rc = put(ln, nullptr, n);
                if (!open_seq) { n_reads += 1; open_seq = true; }
                n_bases += n;
            }
        }
    }
    if (!rc && open_seq) end_seq();
    if (!rc && fill) rc = cs.submit(fill, min_qual_char);
    const int rc2 = cs.close();
    if (R.gz) g_z.close(R.gz); else if (R.f && R.f != stdin) fclose(R.f);
    if (n_reads_out) *n_reads_out = n_reads;
    if (n_bases_out) *n_bases_out = n_bases;
    return rc ? rc : rc2;
}

// Column i of the 32 x 62 binary matrix written into the header.  Any matrix works as long as the records are
// sorted by it; this one is a fixed pseudo-random one.
static uint32_t jf_matrix_column(int i) {
    uint64_t z = 0x6B6D5F62323030ull + (uint64_t)i;          // splitmix64 finaliser (host copy)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)(z >> 17);
}

// A Jellyfish 2.x `binary/sorted` file: "%09d" header length, JSON header (NUL-padded to 8 bytes), then
// ceil(key_len/8)-byte LE key + counter_len-byte LE count per record.  Verified on the five files bundled with
// km (tests/test_jf_writer.py): the records are sorted by pos = M * key over GF(2), masked to
// `size`, where bit b of the key selects column c-1-b of `matrix1`.  Ties (unobserved in the bundled files)
// are broken by key.  The writer emits its own matrix, so readers that binary-search by position stay consistent.
extern "C" int km_table_write_jf(km_table* t, const char* path, uint32_t counter_len) {
    if (!t || !path) return fail(KM_E_ARG, "km_table_write_jf: bad argument");
    if (counter_len == 0) counter_len = 4;
    if (counter_len > 8) return fail(KM_E_ARG, "km_table_write_jf: counter_len must be 1..8");
    std::vector<uint64_t> keys((size_t)t->n_keys);
    std::vector<uint32_t> counts((size_t)t->n_keys);
    uint64_t n = 0;
    if (int rc = km_table_export(t, keys.data(), counts.data(), t->n_keys, &n)) return rc;
    if (n != t->n_keys) return fail(KM_E_ARG, "km_table_write_jf: table holds %llu records, expected %llu", (unsigned long long)n, (unsigned long long)t->n_keys);
    const int kbits = 2 * t->k, kbytes = (kbits + 7) / 8;
    int lsize = 10;
    while (lsize < 32 && (1ull << lsize) < 2 * n) ++lsize;
    const uint64_t size = 1ull << lsize, mask = size - 1;
    // byte-sliced matrix-vector product: tab[j][v] = XOR of the columns selected by byte j of the key
    std::vector<uint32_t> col((size_t)kbits);
    for (int i = 0; i < kbits; ++i) col[(size_t)i] = jf_matrix_column(i);
    std::vector<uint32_t> tab((size_t)8 * 256, 0);
    for (int j = 0; j < 8; ++j)
        for (int v = 0; v < 256; ++v) {
            uint32_t x = 0;
            for (int b = 0; b < 8; ++b) { const int bit = 8 * j + b; if (((v >> b) & 1) && bit < kbits) x ^= col[(size_t)(kbits - 1 - bit)]; }
            tab[(size_t)j * 256 + (size_t)v] = x;
        }
    std::vector<uint64_t> order((size_t)n);      // pos << 32 | index would lose ties on the key: sort indices by (pos, key)
    std::vector<uint32_t> pos((size_t)n);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t x = 0;
        for (int j = 0; j < 8; ++j) x ^= tab[(size_t)j * 256 + ((keys[i] >> (8 * j)) & 0xFF)];
        pos[i] = (uint32_t)(x & mask);
        order[i] = i;
    }
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return pos[a] != pos[b] ? pos[a] < pos[b] : keys[a] < keys[b]; });
    std::string js = "{\"alignment\":8,\"canonical\":";
    js += t->canonical ? "true" : "false";
    js += ",\"cmdline\":[\"km_b200\",\"count\"],\"counter_len\":" + std::to_string(counter_len) + ",\"format\":\"binary/sorted\",\"key_len\":" +
          std::to_string(kbits) + ",\"matrix1\":{\"c\":" + std::to_string(kbits) + ",\"columns\":[";
    for (int i = 0; i < kbits; ++i) { if (i) js += ','; js += std::to_string(col[(size_t)i]); }
    js += "],\"r\":32},\"max_reprobe\":126,\"reprobes\":[1";
    for (int i = 1; i <= 126; ++i) js += "," + std::to_string(i * (i + 1) / 2);
    js += "],\"size\":" + std::to_string(size) + ",\"val_len\":" + std::to_string(8 * counter_len > 12 ? 12 : 8 * counter_len) + "}";
    while ((9 + js.size()) % 8) js += '\0';
    FILE* f = fopen(path, "wb");
    if (!f) return fail(KM_E_IO, "cannot write %s", path);
    char digits[16];
    snprintf(digits, sizeof(digits), "%09zu", js.size());
    bool ok = fwrite(digits, 1, 9, f) == 9 && fwrite(js.data(), 1, js.size(), f) == js.size();
    const size_t rec = (size_t)kbytes + counter_len;
    std::vector<unsigned char> buf;
    buf.reserve(rec << 16);
    const uint64_t cmax = counter_len >= 4 ? 0xFFFFFFFFull : ((1ull << (8 * counter_len)) - 1);
    for (uint64_t i = 0; ok && i < n; ++i) {
        const uint64_t key = keys[order[i]];
        const uint64_t cnt = std::min<uint64_t>(counts[order[i]], cmax);       // a narrow counter saturates
        for (int b = 0; b < kbytes; ++b) buf.push_back((unsigned char)(key >> (8 * b)));
        for (uint32_t b = 0; b < counter_len; ++b) buf.push_back((unsigned char)(b < 8 ? cnt >> (8 * b) : 0));
        if (buf.size() >= (rec << 16)) { ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    if (fclose(f) != 0) ok = false;
    if (!ok) return fail(KM_E_IO, "short write to %s", path);
    return 0;
}
// ---- .jf loader (binary/sorted) ---------------------------------------
static bool json_field(const std::string& js, const char* name, std::string* out) {
    std::string pat = std::string("\"") + name + "\"";
    size_t p = js.find(pat);
    if (p == std::string::npos) return false;
    p = js.find(':', p + pat.size());
    if (p == std::string::npos) return false;
    ++p;
    while (p < js.size() && isspace((unsigned char)js[p])) ++p;
    size_t e = p;
    if (js[p] == '"') { e = js.find('"', p + 1); if (e == std::string::npos) return false; *out = js.substr(p + 1, e - p - 1); return true; }
    while (e < js.size() && js[e] != ',' && js[e] != '}' && !isspace((unsigned char)js[e])) ++e;
    *out = js.substr(p, e - p);
    return true;
}

// records of `rec` bytes (kbytes of key, little-endian, then cbytes of count) decoded and inserted on the device
__global__ void __launch_bounds__(256) km_jf_insert_kernel(TableView T, const uint8_t* __restrict__ raw, uint64_t n_rec, int kbytes,
                                                            int cbytes, unsigned long long* n_new, uint32_t* full) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int rec = kbytes + cbytes;
    unsigned long long mine = 0;
    bool is_full = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += stride) {
        const uint8_t* p = raw + i * (uint64_t)rec;
        uint64_t key = 0, cnt = 0;
        for (int b = 0; b < kbytes; ++b) key |= (uint64_t)p[b] << (8 * b);
        for (int b = 0; b < cbytes; ++b) cnt |= (uint64_t)p[kbytes + b] << (8 * b);
        const int r = table_insert(T, key & T.kmask, cnt > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)cnt, KM_INSERT_OVERWRITE);
        is_full |= r < 0;
        mine += r > 0;
    }
    if (is_full) *full = 1;
    mine = warp_sum64(mine);
    if (warp_leader() && mine) atomicAdd(n_new, mine);
}

// Jellyfish(filename) (km/utils/Jellyfish.py:23-45): the header is parsed on the host, the records -- gigabytes for a real
// sample (example/README.rst:47-48) -- stream through two pinned buffers (pread of the next chunk overlaps the copy and the
// decode + insert kernel of the previous one); nothing is decoded or staged per record on the host.
extern "C" int km_table_open_jf(const char* path, int device, km_table** out) {
    if (!path || !out) return fail(KM_E_ARG, "km_table_open_jf: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(KM_E_IO, "cannot open %s", path);
    char digits[10] = {0};
    if (fread(digits, 1, 9, f) != 9) { fclose(f); return fail(KM_E_IO, "%s: truncated header", path); }
    for (int i = 0; i < 9; ++i) if (!isdigit((unsigned char)digits[i])) { fclose(f); return fail(KM_E_IO, "%s: not a Jellyfish file (no header length)", path); }
    const long hlen = atol(digits);
    std::string js((size_t)hlen, '\0');
    if (fread(&js[0], 1, (size_t)hlen, f) != (size_t)hlen) { fclose(f); return fail(KM_E_IO, "%s: truncated header", path); }
    std::string fmt, canon, key_len, counter_len;
    if (!json_field(js, "format", &fmt) || !json_field(js, "canonical", &canon) || !json_field(js, "key_len", &key_len) ||
        !json_field(js, "counter_len", &counter_len)) { fclose(f); return fail(KM_E_IO, "%s: header lacks format/canonical/key_len/counter_len", path); }
    if (fmt != "binary/sorted") { fclose(f); return fail(KM_E_IO, "%s: unsupported format '%s' (only binary/sorted)", path, fmt.c_str()); }
    const int kbits = atoi(key_len.c_str()), cbytes = atoi(counter_len.c_str());
    if (kbits < 2 || kbits > 62 || (kbits & 1) || cbytes < 1 || cbytes > 8) { fclose(f); return fail(KM_E_IO, "%s: key_len %d / counter_len %d not supported", path, kbits, cbytes); }
    const int kbytes = (kbits + 7) / 8, rec = kbytes + cbytes;
    fseeko(f, 0, SEEK_END);
    const int64_t fsize = (int64_t)ftello(f);
    const int64_t payload = fsize - 9 - hlen;
    if (payload < 0 || payload % rec) { fclose(f); return fail(KM_E_IO, "%s: payload of %lld bytes is not a multiple of %d", path, (long long)payload, rec); }
    const uint64_t n = (uint64_t)(payload / rec);
    km_table* t = nullptr;
    if (int rc = km_table_create(device, kbits / 2, canon == "true", std::max<uint64_t>(n, 1024), &t)) { fclose(f); return rc; }
    // two pinned chunks of whole records
    const size_t chunk_rec = std::max<size_t>(1, std::min<size_t>(((size_t)64 << 20) / (size_t)rec, (size_t)std::max<uint64_t>(n, 1)));
    const size_t chunk_bytes = chunk_rec * (size_t)rec;
    char* pin[2] = {nullptr, nullptr}; char* dev[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    int rc = 0;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) { if (pin[b]) cudaFreeHost(pin[b]); if (dev[b]) cudaFree(dev[b]); if (done[b]) cudaEventDestroy(done[b]); }
        fclose(f);
    };
    for (int b = 0; b < 2 && !rc; ++b) {
        if (cudaMallocHost((void**)&pin[b], chunk_bytes) != cudaSuccess || cudaMalloc((void**)&dev[b], chunk_bytes) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(KM_E_CUDA, "km_table_open_jf: staging buffers of %zu bytes: %s", chunk_bytes, cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc && cudaMemsetAsync(t->d_counter, 0, 16, t->stream) != cudaSuccess) rc = fail(KM_E_CUDA, "memset failed");
    const int fd = fileno(f);
    int64_t at = 9 + hlen;
    int slot = 0;
    for (uint64_t left = n; !rc && left; slot ^= 1) {
        const size_t m = (size_t)std::min<uint64_t>(left, chunk_rec), bytes = m * (size_t)rec;
        if (busy[slot]) { if (cudaEventSynchronize(done[slot]) != cudaSuccess) { rc = fail(KM_E_CUDA, "event sync failed"); break; } busy[slot] = false; }
        size_t got = 0;
        while (got < bytes) {
            const ssize_t r = pread(fd, pin[slot] + got, bytes - got, (off_t)(at + (int64_t)got));
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (got != bytes) { rc = fail(KM_E_IO, "%s: short read", path); break; }
        if (cudaMemcpyAsync(dev[slot], pin[slot], bytes, cudaMemcpyHostToDevice, t->stream) != cudaSuccess) { rc = fail(KM_E_CUDA, "copy failed"); break; }
        km_jf_insert_kernel<<<(int)std::min<uint64_t>((m + 255) / 256, (uint64_t)t->sm_count * 8), 256, 0, t->stream>>>(
            t->view(), (const uint8_t*)dev[slot], (uint64_t)m, kbytes, cbytes, t->d_counter, reinterpret_cast<uint32_t*>(t->d_counter + 1));
        if (cudaGetLastError() != cudaSuccess || cudaEventRecord(done[slot], t->stream) != cudaSuccess) { rc = fail(KM_E_CUDA, "launch failed"); break; }
        busy[slot] = true;
        at += (int64_t)bytes;
        left -= m;
    }
    if (!rc) {
        unsigned long long host[2] = {0, 0};
        if (cudaMemcpyAsync(host, t->d_counter, 16, cudaMemcpyDeviceToHost, t->stream) != cudaSuccess || cudaStreamSynchronize(t->stream) != cudaSuccess)
            rc = fail(KM_E_CUDA, "km_table_open_jf: %s", cudaGetErrorString(cudaGetLastError()));
        else {
            t->n_keys += host[0];
            t->linked = false;
            if ((uint32_t)host[1]) rc = fail(KM_E_FULL, "km_table_open_jf: table full");
        }
    } else cudaStreamSynchronize(t->stream);
    cleanup();
    if (rc) { km_table_close(t); return rc; }
    *out = t;
    return 0;
}
