// ot_writer.cpp -- writes `.ot` files the way tch (the reference's libtorch binding) does, with libtorch itself:
//   VarStore::save  -> Tensor::save_multi -> at_save_multi:  OutputArchive::write(name, tensor) per variable, save_to(path)
//   Tensor::save    -> at_save -> torch::save(tensor, path):  OutputArchive << tensor  (key "0")
// (tch 0.13.0's torch-sys/libtch/torch_api.cpp; tch's source is not vendored in the reference, so these two call
// sequences are restated from its published source.)  This is TEST INFRASTRUCTURE: it gives die_e_b200.nnet.load_ot and
// alphazero.load_training_data a file that was NOT written by the code under test.
//
// usage: ot_writer multi <spec.bin> <out.ot>     spec = u32 n, then per tensor: u32 name_len, name, u32 dtype (0 f32, 1 i8),
//        ot_writer single <spec.bin> <out.ot>           u32 ndim, i64 dims[ndim], raw data
#include <torch/torch.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

static bool read_tensor(std::ifstream &f, std::string &name, torch::Tensor &t) {
    uint32_t nl = 0, dtype = 0, nd = 0;
    if (!f.read(reinterpret_cast<char *>(&nl), 4)) return false;
    name.resize(nl);
    f.read(&name[0], nl);
    f.read(reinterpret_cast<char *>(&dtype), 4);
    f.read(reinterpret_cast<char *>(&nd), 4);
    std::vector<int64_t> dims(nd);
    f.read(reinterpret_cast<char *>(dims.data()), 8 * nd);
    t = torch::empty(dims, dtype == 0 ? torch::kFloat32 : torch::kInt8);
    f.read(reinterpret_cast<char *>(t.data_ptr()), (std::streamsize)t.nbytes());
    return (bool)f;
}

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    const std::string mode = argv[1];
    std::ifstream f(argv[2], std::ios::binary);
    uint32_t n = 0;
    f.read(reinterpret_cast<char *>(&n), 4);
    if (mode == "multi") {
        torch::serialize::OutputArchive archive;
        for (uint32_t i = 0; i < n; ++i) {
            std::string name;
            torch::Tensor t;
            if (!read_tensor(f, name, t)) return 3;
            archive.write(name, t, /*is_buffer=*/false);
        }
        archive.save_to(argv[3]);
    } else {
        std::string name;
        torch::Tensor t;
        if (!read_tensor(f, name, t)) return 3;
        torch::save(t, argv[3]);
    }
    return 0;
}
