/* integration/cnnp_resnet50_bench.c -- BASELINE configs[2] through the REFERENCE'S OWN public API: a ccv_cnnp_model_t ResNet-50 (the
 * v1d variant bin/nnc/imagenet.c trains: three 3x3 stem convolutions 32 / 32 / 64, average-pool-down projection shortcuts,
 * softmax + categorical cross-entropy, nesterov SGD) compiled and fitted with ccv_cnnp_model_compile / ccv_cnnp_model_fit on synthetic
 * data.  Linked against integration/_build/libccv_dropin.so this is the unmodified reference -- model zoo layer, symbolic graph,
 * autograd, memory planner, graph runner, stream contexts -- issuing every command of the step to CCV_NNC_BACKEND_GPU_SM100 (the only
 * GPU backend in that library).  It is a user program of ccv, not part of the backend: what `bin/nnc/imagenet.c` would do without its
 * data pipeline.
 *
 *   cnnp_resnet50_bench --device gpu|cpu --batch N --image S --steps K --warmup W [--classes C] [--lr R] [--seed S]
 *
 * Prints one JSON object: images/s over K timed steps (ccv_nnc_stream_context_wait on both sides), ms/step, the first and last
 * losses (-log p[label] averaged over the batch, from the model's softmax output) and ccv_cnnp_model_memory_size.  --device cpu runs
 * the same program on CPU tensors (the reference's CPU backends): the check of this file that needs no GPU. */
#include <ccv.h>
#include <nnc/ccv_nnc.h>
#include <nnc/ccv_nnc_easy.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* convolution -> batch norm [-> relu] as one sequential model */
static ccv_cnnp_model_t* conv_bn(const int filters, const int k, const int stride, const int no_bias, const int with_relu)
{
	ccv_cnnp_model_t* layers[3];
	int n = 0;
	layers[n++] = ccv_cnnp_convolution(1, filters, DIM_ALLOC(k, k), DIM_ALLOC(), no_bias, HINT((stride, stride), (k / 2, k / 2)), 0, 1, 0);
	layers[n++] = ccv_cnnp_batch_norm(0.9, 1e-4, 1, 0);
	if (with_relu)
		layers[n++] = ccv_cnnp_relu(0);
	return ccv_cnnp_sequential_new(layers, n, 1, 0);
}

/* one bottleneck: 1x1 -> 3x3 (carries the stride) -> 1x1 (x4), summed with the identity or with the projected (average-pooled when
 * strided, then 1x1) input, then relu */
static ccv_cnnp_model_io_t bottleneck(const ccv_cnnp_model_io_t x, const int width, const int stride, const int project)
{
	ccv_cnnp_model_io_t skip = x;
	if (project)
	{
		if (stride > 1)
			skip = ccv_cnnp_model_apply(ccv_cnnp_average_pool(DIM_ALLOC(stride, stride), HINT((stride, stride), (0, 0)), 0), MODEL_IO_LIST(skip));
		skip = ccv_cnnp_model_apply(ccv_cnnp_convolution(1, width * 4, DIM_ALLOC(1, 1), DIM_ALLOC(), 1, HINT((1, 1), (0, 0)), 0, 1, 0), MODEL_IO_LIST(skip));
	}
	ccv_cnnp_model_io_t y = ccv_cnnp_model_apply(conv_bn(width, 1, 1, 0, 1), MODEL_IO_LIST(x));
	y = ccv_cnnp_model_apply(conv_bn(width, 3, stride, 0, 1), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(conv_bn(width * 4, 1, 1, 0, 0), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(ccv_cnnp_sum(0), MODEL_IO_LIST(y, skip));
	return ccv_cnnp_model_apply(ccv_cnnp_relu(0), MODEL_IO_LIST(y));
}

static ccv_cnnp_model_t* resnet50_v1d(const int classes)
{
	static const struct { int width, blocks, stride; } stages[4] = { { 64, 3, 1 }, { 128, 4, 2 }, { 256, 6, 2 }, { 512, 3, 2 } };
	const ccv_cnnp_model_io_t input = ccv_cnnp_input();
	ccv_cnnp_model_io_t y = ccv_cnnp_model_apply(conv_bn(32, 3, 2, 1, 1), MODEL_IO_LIST(input));
	y = ccv_cnnp_model_apply(conv_bn(32, 3, 1, 1, 1), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(conv_bn(64, 3, 1, 1, 1), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(ccv_cnnp_max_pool(DIM_ALLOC(3, 3), HINT((2, 2), (1, 1)), 0), MODEL_IO_LIST(y));
	int s, b;
	for (s = 0; s < 4; s++)
		for (b = 0; b < stages[s].blocks; b++)
			y = bottleneck(y, stages[s].width, b == 0 ? stages[s].stride : 1, b == 0);
	y = ccv_cnnp_model_apply(ccv_cnnp_average_pool(DIM_ALLOC(0, 0), ccv_nnc_no_hint, 0), MODEL_IO_LIST(y)); /* global */
	y = ccv_cnnp_model_apply(ccv_cnnp_flatten(0), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(ccv_cnnp_dense(classes, 0, 0, 1, 0), MODEL_IO_LIST(y));
	y = ccv_cnnp_model_apply(ccv_cnnp_softmax(0), MODEL_IO_LIST(y));
	return ccv_cnnp_model_new(MODEL_IO_LIST(input), MODEL_IO_LIST(y), 1, 0);
}

/* mean of -log p[label] from the model's softmax output (copied to the host) */
static double batch_loss(ccv_nnc_tensor_t* const device_out, ccv_nnc_tensor_t* const host_out, const int* const labels, const int batch, const int classes)
{
	ccv_nnc_cmd_exec(CMD_DATA_TRANSFER_FORWARD(), ccv_nnc_no_hint, 0, TENSOR_LIST(device_out), TENSOR_LIST(host_out), 0);
	double loss = 0;
	int i;
	for (i = 0; i < batch; i++)
	{
		const float p = host_out->data.f32[(size_t)i * classes + labels[i]];
		loss -= log(p > 1e-30f ? p : 1e-30f);
	}
	return loss / batch;
}

int main(int argc, char** argv)
{
	int gpu = 1, batch = 256, image = 224, steps = 10, warmup = 3, classes = 1000, seed = 0, i;
	float lr = 0.01f;
	for (i = 1; i < argc; i++)
	{
		if (!strcmp(argv[i], "--device") && i + 1 < argc) gpu = strcmp(argv[++i], "cpu") != 0;
		else if (!strcmp(argv[i], "--batch") && i + 1 < argc) batch = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--image") && i + 1 < argc) image = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--classes") && i + 1 < argc) classes = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--lr") && i + 1 < argc) lr = (float)atof(argv[++i]);
		else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = atoi(argv[++i]);
		else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
	}
	ccv_nnc_init();
	if (seed) /* the parameter initialisers draw from the calling thread's generator (lib/nnc/ccv_nnc_stream.c:247-281) */
		ccv_nnc_stream_context_set_seed(0, (uint32_t)seed);
	if (gpu && ccv_nnc_device_count(CCV_STREAM_CONTEXT_GPU) < 1)
	{
		printf("{\"unavailable\": \"no GPU device\"}\n");
		return 0;
	}
	ccv_cnnp_model_t* const model = resnet50_v1d(classes);
	/* NHWC activations: the channel-contiguous layout the tensor-core kernels take directly (the reference's cnnp layers follow the input's format) */
	const ccv_nnc_tensor_param_t x_params = gpu ? GPU_TENSOR_NHWC(000, 32F, batch, image, image, 3) : CPU_TENSOR_NHWC(32F, batch, image, image, 3);
	const ccv_nnc_tensor_param_t fit_params = gpu ? GPU_TENSOR_NHWC(000, 32F, batch, classes) : CPU_TENSOR_NHWC(32F, batch, classes);
	/* nesterov SGD with the 1 / batch gradient scale, as bin/nnc/imagenet.c:316 */
	ccv_cnnp_model_compile(model, &x_params, 1, CMD_SGD_FORWARD(1, lr, 1. / batch, 1e-4, 0.9, 0), CMD_CATEGORICAL_CROSSENTROPY_FORWARD());
	ccv_cnnp_model_set_workspace_size(model, 1llu * 1024 * 1024 * 1024);
	/* synthetic batch: U(0, 1) pixels, one-hot labels; built on the host, moved once */
	ccv_nnc_tensor_t* const hx = ccv_nnc_tensor_new(0, CPU_TENSOR_NHWC(32F, batch, image, image, 3), 0);
	ccv_nnc_tensor_t* const hfit = ccv_nnc_tensor_new(0, CPU_TENSOR_NHWC(32F, batch, classes), 0);
	ccv_nnc_tensor_t* const hout = ccv_nnc_tensor_new(0, CPU_TENSOR_NHWC(32F, batch, classes), 0);
	int* const labels = (int*)malloc(sizeof(int) * batch);
	unsigned long long lcg = 88172645463325252ull;
	const size_t nx = (size_t)batch * image * image * 3;
	size_t j;
	for (j = 0; j < nx; j++)
	{
		lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
		hx->data.f32[j] = (float)((lcg >> 40) * (1.0 / 16777216.0));
	}
	memset(hfit->data.f32, 0, sizeof(float) * (size_t)batch * classes);
	for (i = 0; i < batch; i++)
	{
		lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
		labels[i] = (int)((lcg >> 33) % (unsigned)classes);
		hfit->data.f32[(size_t)i * classes + labels[i]] = 1;
	}
	ccv_nnc_tensor_t* const x = ccv_nnc_tensor_new(0, x_params, 0);
	ccv_nnc_tensor_t* const fit = ccv_nnc_tensor_new(0, fit_params, 0);
	ccv_nnc_tensor_t* const out = ccv_nnc_tensor_new(0, fit_params, 0);
	ccv_nnc_cmd_exec(CMD_DATA_TRANSFER_FORWARD(), ccv_nnc_no_hint, 0, TENSOR_LIST(hx, hfit), TENSOR_LIST(x, fit), 0);
	ccv_nnc_stream_context_t* const stream = gpu ? ccv_nnc_stream_context_new(CCV_STREAM_CONTEXT_GPU) : 0;
	double first_loss = 0, last_loss = 0;
	for (i = 0; i < warmup; i++)
	{
		ccv_cnnp_model_fit(model, TENSOR_LIST(x), TENSOR_LIST(fit), TENSOR_LIST(out), 0, stream);
		if (stream)
			ccv_nnc_stream_context_wait(stream);
		if (i == 0)
			first_loss = batch_loss(out, hout, labels, batch, classes);
	}
	const double t0 = now_ms();
	for (i = 0; i < steps; i++)
		ccv_cnnp_model_fit(model, TENSOR_LIST(x), TENSOR_LIST(fit), TENSOR_LIST(out), 0, stream);
	if (stream)
		ccv_nnc_stream_context_wait(stream);
	const double ms = (now_ms() - t0) / (steps > 0 ? steps : 1);
	last_loss = batch_loss(out, hout, labels, batch, classes);
	if (warmup == 0)
		first_loss = last_loss;
	printf("{\"workload\": \"ccv_cnnp_model ResNet-50 v1d fp32 NHWC fwd+bwd+SGD through the reference's own ccv_cnnp_model_fit\", \"device\": \"%s\", \"batch\": %d, \"image\": %d, \"classes\": %d, "
		"\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.4f, \"images_per_sec\": %.2f, \"first_loss\": %.6f, \"last_loss\": %.6f, \"model_memory_bytes\": %llu}\n",
		gpu ? "gpu" : "cpu", batch, image, classes, steps, warmup, ms, batch * 1000.0 / ms, first_loss, last_loss, (unsigned long long)ccv_cnnp_model_memory_size(model));
	fflush(stdout);
	if (stream)
		ccv_nnc_stream_context_free(stream);
	ccv_cnnp_model_free(model);
	ccv_nnc_tensor_free(x), ccv_nnc_tensor_free(fit), ccv_nnc_tensor_free(out);
	ccv_nnc_tensor_free(hx), ccv_nnc_tensor_free(hfit), ccv_nnc_tensor_free(hout);
	free(labels);
	return 0;
}
