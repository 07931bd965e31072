# parity.jl — close the parity loop with the REAL reference on any machine that has Julia >= 1.10, ProtoGrad (this
# repository's /root/reference checkout or the registered package), Zygote 0.7.3, NNlib 0.9 and NPZ.jl:
#
#     julia --project=<env with ProtoGrad, Zygote, NNlib, NPZ> julia/parity.jl [tests/golden] [--device]
#
# For every golden case (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle with fixed
# seeds) it rebuilds the seeded inputs and initial weights, runs the genuine
#     eval_with_pullback(f, m, :Zygote) ; pb() ; step!(state, grad)
# loop of the reference and compares loss, every gradient leaf and the updated parameters with the stored oracle values
# (norms + strided samples).  With --device (needs a B200 and the built libprotograd_b200.so) the same loop is also run
# through ProtoGradB200 (`:B200` backend) and compared with Zygote directly.  This script could not be executed in the
# build image (no Julia); it is the piece SURVEY.md §8(c) asks for so that the "parity unpinned" rows can be closed.
using ProtoGrad, Zygote, NNlib, NPZ, LinearAlgebra, Statistics, Random

const GOLDEN = length(ARGS) >= 1 && !startswith(ARGS[1], "--") ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
const DEVICE = "--device" in ARGS
DEVICE && include(joinpath(@__DIR__, "ProtoGradB200.jl"))

relerr(a, b) = norm(vec(Float64.(a)) .- vec(Float64.(b))) / max(norm(vec(Float64.(b))), eps())

# numpy C-order arrays are Julia arrays with reversed dims: W[in][out] (C) == W[out,in] (Julia), x[N][C][H][W] == x[W,H,C,N]
cdims(a) = permutedims(a, reverse(1:ndims(a)))

"rebuild the model of a golden case from its stored initial leaves (`p0_<i>` entries, vec(m) order)"
function build(case, d)
    leaf(i) = cdims(d["p0_$(i)"])
    if case == "c1_mlp"
        return Compose(Linear(leaf(0), leaf(1)), x -> relu.(x), Linear(leaf(2), leaf(3)), softmax)
    elseif case == "c3_conv" || case == "c5_finetune"
        # c3 head (tests/golden/make_golden.py C3_SPEC, examples/02_mnist_conv.jl:36-42): 256 -> 120 -> relu -> 84 -> relu -> 10;
        # c5 head (C5_SPEC): one trainable Linear 256 -> 10 — layer 8 of the Compose in both cases
        head = case == "c3_conv" ?
            (Linear(leaf(4), leaf(5)), x -> relu.(x), Linear(leaf(6), leaf(7)), x -> relu.(x), Linear(leaf(8), leaf(9))) :
            (Linear(leaf(4), leaf(5)),)
        return Compose(Conv(leaf(0), reshape(leaf(1), 1, 1, :)), x -> relu.(x), x -> maxpool(x, (2, 2)),
                       Conv(leaf(2), reshape(leaf(3), 1, 1, :)), x -> relu.(x), x -> maxpool(x, (2, 2)),
                       x -> reshape(x, :, size(x, 4)), head..., softmax)
    end
    error("unknown case $case")
end

function run_case(case)
    d = npzread(joinpath(GOLDEN, case * ".npz"))
    haskey(d, "x") || (println("$case: fixture holds no raw inputs (regenerate with python -m tests.golden.make_golden)"); return true)
    x, lab = cdims(d["x"]), cdims(d["label"])
    full = build(case, d)
    # c5: frozen conv trunk captured by the closure, only the Linear head is differentiated (test/test_fine_tuning.jl:29-33)
    frozen = case == "c5_finetune"
    m = frozen ? full.layers[8] : full
    first_leaf = frozen ? 4 : 0                                   # index of m's first leaf in the fixture's vec(m) order
    objective = frozen ? (h -> cross_entropy(softmax(h(Compose(full.layers[1:7]...)(x))), lab)) : (mm -> cross_entropy(mm(x), lab))
    opt = case == "c1_mlp" ? Adam(stepsize = 1e-3) : case == "c3_conv" ? NesterovMethod(stepsize = 1e-2) : GradientDescent(stepsize = 0.1)
    state = init(opt, m)
    ok = true
    for step in 1:length(d["losses"])
        val, pb = eval_with_pullback(objective, m, :Zygote)
        grad = pb()
        if step == 1
            e = abs(val - d["loss"]) / abs(d["loss"])
            println("$case: loss $(val) vs oracle $(d["loss"])  rel $(e)")
            ok &= e <= 1e-5
            for (i, g) in enumerate(ProtoGrad.allparams(grad))
                k = first_leaf + i - 1
                n = norm(vec(Float64.(g)))
                e = abs(n - d["grad$(k)_norm"]) / max(d["grad$(k)_norm"], eps())
                println("   grad leaf $(k): norm $(n) vs $(d["grad$(k)_norm"])  rel $(e)")
                ok &= e <= 1e-5
            end
        end
        e = abs(val - d["losses"][step]) / abs(d["losses"][step])
        ok &= e <= 2e-5
        step!(state, grad)
    end
    for (i, p) in enumerate(ProtoGrad.allparams(m))
        k = first_leaf + i - 1
        n = norm(vec(Float64.(p)))
        e = abs(n - d["param$(k)_norm"]) / max(d["param$(k)_norm"], eps())
        println("   param leaf $(k) after $(length(d["losses"])) steps: norm rel $(e)")
        ok &= e <= 1e-5
    end
    if DEVICE
        B = Main.ProtoGradB200
        m2 = frozen ? build(case, d).layers[8] : build(case, d)
        dm = B.to_device(m2)
        val_d, pb_d = eval_with_pullback(objective, dm, :B200)
        val_z, pb_z = eval_with_pullback(objective, m2, :Zygote)
        gd, gz = B.to_host(pb_d()), pb_z()
        println("$case: :B200 loss $(val_d) vs :Zygote $(val_z); gradient rel ", relerr(vec(gd), vec(gz)))
        ok &= abs(val_d - val_z) <= 1e-5 * abs(val_z) && relerr(vec(gd), vec(gz)) <= 1e-5
    end
    return ok
end

all_ok = all(run_case(c) for c in ("c1_mlp", "c3_conv", "c5_finetune"))
println(all_ok ? "PARITY OK" : "PARITY FAILED")
exit(all_ok ? 0 : 1)
