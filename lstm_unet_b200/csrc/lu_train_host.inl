// Host side of the training step (included by lu_api.cu).
static void train_layout(lu_handle_s* h, size_t& off) {
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  h->tr.off_loss_acc = take(64);
}
static void train_destroy(lu_handle_s*) {}

static int train_loss_backward(lu_handle_s* h, const float*, const float*, float*, float*, void*) {
  LU_REQUIRE(h, "null handle");
  LU_FAIL("lu_loss_backward: backward pass not built yet");
}

static int train_adam(lu_handle_s* h, const float* g, float* m, float* v, float lr, float b1, float b2, float eps,
                      int64_t step, void* stream) {
  LU_REQUIRE(h && h->dparams && g && m && v, "null argument");
  LU_REQUIRE(step >= 1, "Adam step is 1-based");
  LuAdam a;
  a.p = h->dparams; a.g = g; a.m = m; a.v = v; a.b1 = b1; a.b2 = b2; a.eps = eps;
  a.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step)));
  pf(h, h->n_train, stream, a);
  h->packed = false;
  return 0;
}
