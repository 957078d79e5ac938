// Host side of mds_post_processing (included by mds_api.cu).
extern "C" size_t mds_post_processing_workspace_bytes(int n_frames, int num_classes) {
    if (n_frames <= 0 || num_classes <= 0) return 0;
    return al256((size_t)n_frames * num_classes * 4) + al256((size_t)n_frames * num_classes) + 256;
}

extern "C" int mds_post_processing(const float* raw, int n_frames, int num_classes, const double* weights_host, int radius,
                                   float height, int distance, int* out_index, float* out_conf, int* out_count, void* ws,
                                   size_t ws_bytes, void* stream) {
    if (!raw || !weights_host || !out_index || !out_conf || !out_count || !ws) return fail(MDS_ERR_INVALID, "post_processing: null argument");
    if (n_frames <= 0 || num_classes <= 0 || num_classes > 65535) return fail(MDS_ERR_INVALID, "post_processing: bad sizes");
    if (radius < 0 || radius > kPostMaxRadius) return fail(MDS_ERR_INVALID, "post_processing: radius %d exceeds %d (sigma too large)", radius, kPostMaxRadius);
    if (distance < 1) return fail(MDS_ERR_INVALID, "post_processing: `distance` must be greater or equal to 1");    // scipy's message
    if (ws_bytes < mds_post_processing_workspace_bytes(n_frames, num_classes)) return fail(MDS_ERR_WORKSPACE, "post_processing: workspace too small");
    PostParams p;
    memset(&p, 0, sizeof(p));
    p.raw = raw; p.N = n_frames; p.K = num_classes; p.radius = radius; p.distance = distance; p.height = height;
    for (int i = 0; i < 2 * radius + 1; ++i) p.w[i] = weights_host[i];
    Arena ar(ws, ws_bytes);
    p.smooth = ar.take<float>((size_t)n_frames * num_classes);
    p.state = ar.take<unsigned char>((size_t)n_frames * num_classes);
    p.out_index = out_index; p.out_conf = out_conf; p.out_count = out_count;
    post_processing_kernel<<<num_classes, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    LAUNCH_CHECK("post_processing");
    return MDS_OK;
}
