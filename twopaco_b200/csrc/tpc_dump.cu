// tpc_dump.cu -- the consumer side of de_bruijn.bin on the GPU:
//   * `graphdump -f seq`   (graphdump.cpp:160-168): one line "chr pos id" per record, in file order
//   * `graphdump -f group` (graphdump.cpp:120-158): occurrences grouped by (signed) id, every class sorted by
//     (chr, pos), classes sorted by their first occurrence, one line "chr pos; chr pos; ..." per class
//   * canonical relabelling: ids renumbered 1.. by first appearance, first occurrence
//     positive -- the parity definition of SURVEY 8(c), here at sizes (C2: 1.5e8, C3: 2.5e8 records) that a host
//     canonicaliser does not handle in reasonable time.
// Reader semantics are JunctionPositionReader's (junctionapi.h:76-99): a unit is a separator when pos == 0xFFFFFFFF
// or id == INT64_MAX; every separator advances the sequence index.
// The text is produced on the device (digit counts -> prefix sums -> one thread per record writes its characters);
// sorting is a stable library radix sort (cub) by id, so file order == (chr, pos) order survives inside a class.
#include "tpc_dump_common.cuh"

extern "C" {

int tpc_graphdump_device(const uint8_t* dev_image, uint64_t image_bytes, uint32_t format, void* stream, uint8_t** dev_text,
                         uint64_t* text_bytes) {
    if (!dev_text || !text_bytes || (image_bytes && !dev_image)) return set_error("null argument");
    if (format > 2) return set_error("format must be 0 (seq), 1 (group) or 2 (dot)");
    if ((uintptr_t)dev_image & 3) return set_error("the image must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st);
    const uint32_t* img = reinterpret_cast<const uint32_t*>(dev_image);
    const uint64_t n_units = image_bytes / 12;   // a truncated trailing unit is ignored, as the reader does (junctionapi.h:88-91)
    Records R;
    if (int rc = load_records(sc, img, n_units, &R)) return rc;
    const uint64_t m = R.m;
    unsigned long long *len = nullptr, *off = nullptr;
    CKD(sc.alloc(&len, m + 1));
    CKD(sc.alloc(&off, m + 1));
    CKD(cudaMemsetAsync(len + m, 0, 8, st));
    char* text = nullptr;
    unsigned long long total = 0;
    if (format == 2) {
        const char head[] = "digraph G\n{\n\trankdir = LR\n", tail[] = "}\n";
        const uint64_t nh = sizeof head - 1, nt = sizeof tail - 1;
        if (m) k_dot_len<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, len);
        if (int rc = exclusive_sum(sc, len, off, m + 1)) return rc;
        CKD(cudaMemcpyAsync(&total, off + m, 8, cudaMemcpyDeviceToHost, st));
        CKD(cudaStreamSynchronize(st));
        CKD(cudaMallocAsync((void**)&text, total + nh + nt + 16, st));
        CKD(cudaMemcpyAsync(text, head, nh, cudaMemcpyHostToDevice, st));
        if (m) k_dot_write<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, off, text + nh);
        CKD(cudaMemcpyAsync(text + nh + total, tail, nt, cudaMemcpyHostToDevice, st));
        total += nh + nt;
    } else if (format == 0) {
        if (m) k_seq_len<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, len);
        if (int rc = exclusive_sum(sc, len, off, m + 1)) return rc;
        CKD(cudaMemcpyAsync(&total, off + m, 8, cudaMemcpyDeviceToHost, st));
        CKD(cudaStreamSynchronize(st));
        CKD(cudaMallocAsync((void**)&text, std::max<unsigned long long>(total, 16), st));
        if (m) k_seq_write<<<grid_for(m), 256, 0, st>>>(R.chr, R.pos, R.id, m, off, text);
    } else {
        unsigned long long* key = nullptr;
        CKD(sc.alloc(&key, m));
        if (m) k_signed_keys<<<grid_for(m), 256, 0, st>>>(R.id, m, key);
        Classes C;
        if (int rc = build_classes(sc, key, m, &C)) return rc;
        if (m) k_group_len<<<grid_for(m), 256, 0, st>>>(C.idx, R.chr, R.pos, m, len);
        if (int rc = exclusive_sum(sc, len, off, m + 1)) return rc;       // member text offsets in id-sorted order
        unsigned long long members_text = 0;
        CKD(cudaMemcpyAsync(&members_text, off + m, 8, cudaMemcpyDeviceToHost, st));
        CKD(cudaStreamSynchronize(st));
        total = members_text + C.groups;
        unsigned long long *clen = nullptr, *cbase_ranked = nullptr, *cbase = nullptr;
        CKD(sc.alloc(&clen, C.groups));
        CKD(sc.alloc(&cbase_ranked, C.groups));
        CKD(sc.alloc(&cbase, C.groups));
        CKD(cudaMallocAsync((void**)&text, std::max<unsigned long long>(total, 16), st));
        if (C.groups) {
            k_class_len<<<grid_for(C.groups), 256, 0, st>>>(C.order, C.first_e, off, C.groups, m, members_text, clen);
            if (int rc = exclusive_sum(sc, clen, cbase_ranked, C.groups)) { cudaFreeAsync(text, st); return rc; }
            k_class_base<<<grid_for(C.groups), 256, 0, st>>>(C.order, cbase_ranked, C.groups, cbase);
            k_group_write<<<grid_for(m), 256, 0, st>>>(C.idx, C.head, C.gid_incl, C.first_e, off, cbase, R.chr, R.pos, m, text);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        if (text) cudaFreeAsync(text, st);
        return set_error("CUDA error %s in graphdump (%s)", cudaGetErrorName(e), cudaGetErrorString(e));
    }
    *dev_text = (uint8_t*)text;
    *text_bytes = total;
    return 0;
}

int tpc_canonical_image_device(const uint8_t* dev_image, uint64_t image_bytes, void* stream, uint8_t* dev_out, uint64_t* n_classes) {
    if ((image_bytes && (!dev_image || !dev_out))) return set_error("null argument");
    if (((uintptr_t)dev_image | (uintptr_t)dev_out) & 3) return set_error("images must be 4-byte aligned");
    if (image_bytes % 12) return set_error("image size is not a multiple of 12");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc(st);
    const uint32_t* img = reinterpret_cast<const uint32_t*>(dev_image);
    const uint64_t n_units = image_bytes / 12;
    Records R;
    if (int rc = load_records(sc, img, n_units, &R)) return rc;
    const uint64_t m = R.m;
    unsigned long long* key = nullptr;
    CKD(sc.alloc(&key, m));
    if (m) k_abs_keys<<<grid_for(m), 256, 0, st>>>(R.id, m, key);
    Classes C;
    if (int rc = build_classes(sc, key, m, &C)) return rc;
    uint32_t* rank = nullptr;
    long long* canon = nullptr;
    CKD(sc.alloc(&rank, C.groups));
    CKD(sc.alloc(&canon, m));
    if (C.groups) {
        k_class_rank<<<grid_for(C.groups), 256, 0, st>>>(C.order, C.groups, rank);
        k_canon_ids<<<grid_for(m), 256, 0, st>>>(C.idx, C.gid_incl, C.first_idx, rank, R.id, m, canon);
    }
    if (n_units) k_canon_image<<<grid_for(n_units), 256, 0, st>>>(img, n_units, R.sep_before, canon, reinterpret_cast<uint32_t*>(dev_out));
    CKD(cudaGetLastError());
    CKD(cudaStreamSynchronize(st));
    if (n_classes) *n_classes = C.groups;
    return 0;
}

// file -> text file (NULL / "-" = stdout): what `graphdump -f seq|group <file>` prints
int tpc_graphdump_file(const char* image_path, const char* format, const char* out_path) {
    if (!image_path || !format) return set_error("null argument");
    const uint32_t fmt = !strcmp(format, "seq") ? 0u : !strcmp(format, "group") ? 1u : !strcmp(format, "dot") ? 2u : 3u;
    if (fmt > 2) return set_error("only the 'seq', 'group' and 'dot' output formats are produced on the GPU");
    FILE* f = fopen(image_path, "rb");
    if (!f) return set_error("Can't open file %s", image_path);
    fseek(f, 0, SEEK_END);
    const uint64_t bytes = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fclose(f);
        return set_error("no CUDA device: twopaco_b200 has no CPU fallback");
    }
    uint8_t* d_img = nullptr;
    void* h = nullptr;
    int rc = 0;
    const uint64_t kPiece = 64ull << 20;
    if (cudaMalloc(&d_img, std::max<uint64_t>(bytes, 16)) != cudaSuccess || cudaMallocHost(&h, std::min<uint64_t>(std::max<uint64_t>(bytes, 16), kPiece)) != cudaSuccess)
        rc = set_error("out of device / pinned memory for a %llu-byte image", (unsigned long long)bytes);
    for (uint64_t lo = 0; rc == 0 && lo < bytes; lo += kPiece) {
        const uint64_t n = std::min(kPiece, bytes - lo);
        if (fread(h, 1, n, f) != n) rc = set_error("Can't read file %s", image_path);
        else if (cudaMemcpy(d_img + lo, h, n, cudaMemcpyHostToDevice) != cudaSuccess) rc = set_error("host to device copy failed");
    }
    fclose(f);
    uint8_t* d_text = nullptr;
    uint64_t tbytes = 0;
    if (rc == 0) rc = tpc_graphdump_device(d_img, bytes, fmt, nullptr, &d_text, &tbytes);
    if (rc == 0 && std::min<uint64_t>(tbytes, kPiece) > std::min<uint64_t>(std::max<uint64_t>(bytes, 16), kPiece)) {
        cudaFreeHost(h);   // the text is larger than the image: a larger staging buffer
        h = nullptr;
        if (cudaMallocHost(&h, std::min<uint64_t>(tbytes, kPiece)) != cudaSuccess) rc = set_error("out of pinned host memory");
    }
    if (rc == 0) {
        FILE* o = (!out_path || !strcmp(out_path, "-")) ? stdout : fopen(out_path, "wb");
        if (!o) rc = set_error("Can't create the output file");
        for (uint64_t lo = 0; rc == 0 && lo < tbytes; lo += kPiece) {
            const uint64_t n = std::min(kPiece, tbytes - lo);
            if (cudaMemcpy(h, d_text + lo, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error("device to host copy failed");
            else if (fwrite(h, 1, n, o) != n) rc = set_error("Can't write to the output file");
        }
        if (o && o != stdout && fclose(o) != 0 && rc == 0) rc = set_error("Can't write to the output file");
        if (o == stdout) fflush(stdout);
    }
    if (d_text) cudaFree(d_text);
    if (d_img) cudaFree(d_img);
    if (h) cudaFreeHost(h);
    return rc;
}

}  // extern "C"
