/* JNI shim between epfl.distributed.nativ.DsgdNative (Scala, see DsgdNative.scala) and the C ABI of
 * include/dsgd.h.  Compile-gated: the build image has no JDK (no jni.h); on a box with one:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libdsgd_jni.so dsgd_jni.c -L.. -ldsgd
 * Arrays are pinned with Get/ReleasePrimitiveArrayCritical: the C ABI consumes host buffers before returning.
 * tests/test_abi_surface.py compiles this file against a minimal stand-in jni.h (tests/jni_mock/) -- a syntax and type
 * check against include/dsgd.h, not a run under a JVM. */
#ifdef DSGD_HAVE_JNI
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "dsgd.h"

#define CTX(h) ((dsgd_ctx *)(intptr_t)(h))
#define PIN(env, arr) ((arr) ? (*(env))->GetPrimitiveArrayCritical((env), (arr), NULL) : NULL)
#define UNPIN(env, arr, p, mode) do { if (arr) (*(env))->ReleasePrimitiveArrayCritical((env), (arr), (p), (mode)); } while (0)

JNIEXPORT jlong JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_create(JNIEnv *env, jobject self, jint device, jint dim,
                                                                          jdouble lambda, jint rank, jint world, jint flags) {
  dsgd_ctx *ctx = NULL;
  int rc = dsgd_create(&ctx, device, dim, lambda, rank, world, (uint32_t)flags);
  return rc == DSGD_OK ? (jlong)(intptr_t)ctx : (jlong)rc; /* negative = error code */
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_destroy(JNIEnv *env, jobject self, jlong h) {
  return dsgd_destroy(CTX(h));
}

JNIEXPORT jstring JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_lastError(JNIEnv *env, jobject self, jlong h) {
  return (*env)->NewStringUTF(env, dsgd_last_error(CTX(h)));
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_gradient(JNIEnv *env, jobject self, jlong h, jdoubleArray w,
                                                                           jintArray samples, jdoubleArray grad) {
  const jsize n = (*env)->GetArrayLength(env, samples);
  double *pw = PIN(env, w); int32_t *ps = PIN(env, samples); double *pg = PIN(env, grad);
  int rc = dsgd_gradient(CTX(h), pw, ps, n, pg, NULL);   /* SlaveImpl.gradient, core/Slave.scala:142-157 */
  UNPIN(env, grad, pg, 0); UNPIN(env, samples, ps, JNI_ABORT); UNPIN(env, w, pw, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_forward(JNIEnv *env, jobject self, jlong h, jdoubleArray w,
                                                                          jintArray samples, jdoubleArray preds) {
  const jsize n = (*env)->GetArrayLength(env, samples);
  double *pw = PIN(env, w); int32_t *ps = PIN(env, samples); double *pp = PIN(env, preds);
  int rc = dsgd_forward(CTX(h), pw, ps, n, pp);          /* SlaveImpl.forward, core/Slave.scala:129-140 */
  UNPIN(env, preds, pp, 0); UNPIN(env, samples, ps, JNI_ABORT); UNPIN(env, w, pw, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_syncSteps(JNIEnv *env, jobject self, jlong h, jintArray samples,
                                                                            jlong n_per_step, jlong n_steps, jdouble lr,
                                                                            jdoubleArray losses) {
  int32_t *ps = PIN(env, samples); double *pl = PIN(env, losses);
  int rc = dsgd_sync_steps(CTX(h), ps, n_per_step, n_steps, lr, pl);   /* Master.fit's batch loop, core/Master.scala:179-198 */
  UNPIN(env, losses, pl, 0); UNPIN(env, samples, ps, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_loadCsr(JNIEnv *env, jobject self, jlong h, jlongArray rowPtr,
                                                                          jintArray col, jfloatArray value, jbyteArray label) {
  const jsize n_rows = (*env)->GetArrayLength(env, rowPtr) - 1;   /* Dataset.rcv1 rows as CSR (utils/Dataset.scala:23-47) */
  const jsize nnz = (*env)->GetArrayLength(env, col);
  int64_t *rp = PIN(env, rowPtr); int32_t *pc = PIN(env, col); float *pv = PIN(env, value); int8_t *pl = PIN(env, label);
  int rc = dsgd_load_csr(CTX(h), n_rows, nnz, rp, pc, pv, pl);
  UNPIN(env, label, pl, JNI_ABORT); UNPIN(env, value, pv, JNI_ABORT); UNPIN(env, col, pc, JNI_ABORT); UNPIN(env, rowPtr, rp, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_computeDimSparsity(JNIEnv *env, jobject self, jlong h,
                                                                                     jlong nTrain, jdoubleArray out) {
  double *po = PIN(env, out);
  int rc = dsgd_compute_dim_sparsity(CTX(h), nTrain, po);        /* Main.scala:54-65 */
  UNPIN(env, out, po, 0);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_setWeights(JNIEnv *env, jobject self, jlong h, jdoubleArray w) {
  double *pw = PIN(env, w);
  int rc = dsgd_set_weights(CTX(h), pw);
  UNPIN(env, w, pw, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_getWeights(JNIEnv *env, jobject self, jlong h, jdoubleArray w) {
  double *pw = PIN(env, w);
  int rc = dsgd_get_weights(CTX(h), pw);
  UNPIN(env, w, pw, 0);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_eval(JNIEnv *env, jobject self, jlong h, jdoubleArray w,
                                                                       jlong rowBegin, jlong rowEnd, jdoubleArray lossAcc) {
  double *pw = PIN(env, w); double *pla = PIN(env, lossAcc);     /* lossAcc(0) = loss, lossAcc(1) = accuracy */
  int rc = dsgd_eval(CTX(h), pw, rowBegin, rowEnd, pla, pla + 1); /* Master.localLoss / localAccuracy, core/Master.scala:100-107 */
  UNPIN(env, lossAcc, pla, 0); UNPIN(env, w, pw, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_startAsync(JNIEnv *env, jobject self, jlong h, jdoubleArray w0,
                                                                             jintArray assigned, jint batch, jdouble lr,
                                                                             jint concurrency, jlong maxUpdates, jlong seed) {
  const jsize n = (*env)->GetArrayLength(env, assigned);
  double *pw = PIN(env, w0); int32_t *pa = PIN(env, assigned);
  int rc = dsgd_start_async(CTX(h), pw, pa, n, batch, lr, concurrency, maxUpdates, (uint64_t)seed);  /* core/Slave.scala:159-175 */
  UNPIN(env, assigned, pa, JNI_ABORT); UNPIN(env, w0, pw, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_stopAsync(JNIEnv *env, jobject self, jlong h) {
  return dsgd_stop_async(CTX(h));                                /* core/Slave.scala:187-195 */
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_updateGrad(JNIEnv *env, jobject self, jlong h, jintArray idx,
                                                                             jdoubleArray value) {
  const jsize n = (*env)->GetArrayLength(env, idx);
  int32_t *pi = PIN(env, idx); double *pv = PIN(env, value);
  int rc = dsgd_update_grad(CTX(h), pi, pv, n);                  /* core/Slave.scala:177-185 */
  UNPIN(env, value, pv, JNI_ABORT); UNPIN(env, idx, pi, JNI_ABORT);
  return rc;
}

JNIEXPORT jint JNICALL Java_epfl_distributed_nativ_DsgdNative_00024_asyncUpdates(JNIEnv *env, jobject self, jlong h, jlongArray out) {
  int64_t *po = PIN(env, out);
  int rc = dsgd_async_updates(CTX(h), po);                       /* GradState.updates, core/ml/GradState.scala:8 */
  UNPIN(env, out, po, 0);
  return rc;
}
#endif /* DSGD_HAVE_JNI */
