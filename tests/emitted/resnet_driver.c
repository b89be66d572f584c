/* resnet_driver.c -- minimal driver for an ACE-emitted CIFAR ResNet translation unit.
 *
 * Plays the role of the reference's resnet_cifar.main.inc
 * (fhe-cmplr/rtlib/ant/dataset/resnet_cifar.main.inc:35-125: Prepare_context, Prepare_input,
 * Run_main_graph, Handle_output, Finalize_context) without the CIFAR reader: the image is the
 * synthetic one of SURVEY.md 8(d) config 1 (LCG x = x*1664525 + 1013904223, seed 12345 + image
 * index, mapped to [-0.5, 0.5)).  The emitted model is #included unmodified:
 *     cc -DMODEL_INC='"<path>/resnet20_cifar10_pre.onnx.inc"' -I include resnet_driver.c -lace_b200
 * The same file also builds against the reference's own headers and rtlib (CPU baseline).
 * usage: resnet_driver [n_images] [n_classes]
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "common/rtlib.h"
#include MODEL_INC

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char** argv) {
  int n_images  = argc > 1 ? atoi(argv[1]) : 1;
  int n_classes = argc > 2 ? atoi(argv[2]) : 10;
  double t0 = now_s();
  Prepare_context();
  double t_ctx = now_s() - t0;
  printf("[driver] Prepare_context %.3f s\n", t_ctx);
  for (int img = 0; img < n_images; img++) {
    TENSOR*  in = Alloc_tensor(1, 3, 32, 32, NULL);
    uint32_t x  = 12345u + (uint32_t)img;
    for (size_t i = 0; i < TENSOR_SIZE(in); i++) {
      x = x * 1664525u + 1013904223u;
      in->_vals[i] = (double)(x >> 8) / 16777216.0 - 0.5;
    }
    double t1 = now_s();
    Prepare_input(in, "input");
    Free_tensor(in);
    double t2 = now_s();
    Run_main_graph();
    double  t3  = now_s();
    double* out = Handle_output("output");
    double  t4  = now_s();
    printf("[driver] image %d: encrypt %.3f s, Main_graph %.3f s, decrypt %.3f s\n", img, t2 - t1,
           t3 - t2, t4 - t3);
    printf("[driver] logits %d:", img);
    for (int k = 0; k < n_classes; k++) printf(" %.17g", out[k]);
    printf("\n");
    free(out);
  }
  Finalize_context();
  return 0;
}
