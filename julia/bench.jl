# bench.jl — the reference CPU path timed with the genuine package (Julia >= 1.10, ProtoGrad, Zygote, NNlib), for the
# configuration bench.py quotes (BASELINE configs[3]: wide MLP 784-4096-4096-10, softmax + cross-entropy, Adam,
# 8,192-sample steps).  Prints one JSON line shaped like `bench.py --impl reference`.  With --device it times the same
# loop through ProtoGradB200 (`:B200` backend).  Not executable in the build image (no Julia).
#
#     julia --project=<env> -t auto julia/bench.jl [--steps K] [--warmup W] [--batch B] [--device]
using ProtoGrad, Zygote, NNlib, LinearAlgebra, Random, Statistics

argval(name, default) = (i = findfirst(==(name), ARGS); i === nothing ? default : parse(Int, ARGS[i+1]))
const STEPS, WARMUP, BATCH = argval("--steps", 5), argval("--warmup", 1), argval("--batch", 8192)
const DEVICE = "--device" in ARGS
DEVICE && include(joinpath(@__DIR__, "ProtoGradB200.jl"))

Random.seed!(0)
model = Compose(Linear(784 => 4096), x -> relu.(x), Linear(4096 => 4096), x -> relu.(x), Linear(4096 => 10), softmax)
x = rand(Float32, 784, BATCH)
label = zeros(Float32, 10, BATCH)
for j in 1:BATCH
    label[rand(1:10), j] = 1
end
m = DEVICE ? Main.ProtoGradB200.to_device(model) : model
backend = DEVICE ? :B200 : :Zygote
state = init(Adam(stepsize = 1e-3), m)
function step()
    val, pb = eval_with_pullback(mm -> cross_entropy(mm(x), label), m, backend)
    step!(state, pb())
    return val
end
for _ in 1:WARMUP
    step()
end
DEVICE && Main.ProtoGradB200.sync()
t0 = time()
for _ in 1:STEPS
    step()
end
DEVICE && Main.ProtoGradB200.sync()
dt = time() - t0
v = BATCH * STEPS / dt
println("""{"impl": "$(DEVICE ? "b200-julia" : "reference")", "metric": "train samples/sec per step", "value": $v, "unit": "samples/s", "steps": $STEPS, "warmup": $WARMUP, "ms_per_step": $(1e3 * dt / STEPS), "higher_is_better": true, "dtype": "f32", "data": "synthetic", "config": {"workload": "wide MLP 784-4096-4096-10 (BASELINE configs[3]), softmax+cross-entropy, Adam(1e-3), $BATCH samples/step"}, "cpu_baseline": {"value": $v, "unit": "samples/s", "cores": $(Sys.CPU_THREADS), "kind": "reference", "sample": "genuine ProtoGrad.jl + Zygote + NNlib, $(Threads.nthreads()) Julia threads, BLAS threads $(BLAS.get_num_threads())"}}""")
