// placeholder
struct sedb_cnn { int dummy; };
struct sedb_m5 { int dummy; };
static int sedb_cnn_kernels_init() { return 0; }
static int sedb_cnn_forward_pipeline(sedb_ctx_t*, sedb_cnn_t*, long long, long long, float*) { return fail("cnn not built"); }
extern "C" {
int sedb_cnn_create(sedb_ctx_t*, const int*, const int*, int, int, sedb_cnn_t**) { return fail("nyi"); }
int sedb_cnn_destroy(sedb_cnn_t*) { return 0; }
int sedb_cnn_load(sedb_cnn_t*, const float* const*, int, void*) { return fail("nyi"); }
long long sedb_cnn_out_frames(const sedb_cnn_t*, long long) { return 0; }
size_t sedb_cnn_workspace_bytes(const sedb_cnn_t*, long long, long long) { return 0; }
int sedb_cnn_forward(sedb_cnn_t*, const float*, long long, long long, float*, float*, void*, size_t, void*) { return fail("nyi"); }
int sedb_m5_create(sedb_ctx_t*, int, sedb_m5_t**) { return fail("nyi"); }
int sedb_m5_destroy(sedb_m5_t*) { return 0; }
int sedb_m5_load(sedb_m5_t*, const float* const*, int, void*) { return fail("nyi"); }
size_t sedb_m5_workspace_bytes(const sedb_m5_t*, long long) { return 0; }
int sedb_m5_forward(sedb_m5_t*, const float*, long long, float*, void*, size_t, void*) { return fail("nyi"); }
}
