/* JNI shim between epfl.distributed.nativ.DsgdNative (Scala, see DsgdNative.scala) and the C ABI of
 * include/dsgd.h.  Compile-gated: the build image has no JDK (no jni.h); on a box with one:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libdsgd_jni.so dsgd_jni.c -L.. -ldsgd
 * Arrays are pinned with Get/ReleasePrimitiveArrayCritical: the C ABI consumes host buffers before returning. */
#ifdef DSGD_HAVE_JNI
#include <jni.h>
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
/* loadCsr, computeDimSparsity, set/getWeights, eval, startAsync, stopAsync, updateGrad, asyncUpdates follow the
 * same pin -> call -> unpin pattern, one ABI call each. */
#endif /* DSGD_HAVE_JNI */
