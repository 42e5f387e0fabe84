// smallk_b200 host — the random source the clustering drivers take by reference.
// Same class name and member functions as the reference's common/include/random.hpp:22-200 (std::mt19937 behind
// uniform distributions), so Clust/ClustSparse keep their signatures and a given seed produces the same
// initialiser stream as the reference's sequential generator (matrix_generator.hpp:229-248).
#pragma once

#include <ctime>
#include <random>
#include <sstream>
#include <string>

class Random
{
public:
    Random() {}

    void SeedFromTime() { engine_.seed(static_cast<unsigned long>(time(0))); }
    void SeedFromRandomDevice() { std::random_device rd; engine_.seed(rd()); }
    void SeedFromInt(const int s) { engine_.seed(s); }
    void SetDefaultState() { engine_.seed(); }
    std::string GetState() { std::stringstream s; s << engine_; return s.str(); }
    void SetState(const std::string& state) { std::stringstream s(state); s >> engine_; }

    int RandomInt() { return dist_int_(engine_); }          // [0, INT_MAX]
    float RandomFloat() { return dist_float_(engine_); }    // [0, 1)
    double RandomDouble() { return dist_double_(engine_); } // [0, 1)

    int RandomRangeInt(const int& a, const int& b) { return a + (RandomInt() % (b - a)); }
    float RandomRangeFloat(const float& a, const float& b) { return a + (b - a) * RandomFloat(); }
    double RandomRangeDouble(const double& a, const double& b) { return a + (b - a) * RandomDouble(); }

    // (center - radius, center + radius)
    int RandomInt(const int& center, const int& radius) { return center + (RandomInt() % (2 * radius)) - radius; }
    float RandomFloat(const float& center, const float& radius) { return center + 2.0f * radius * RandomFloat() - radius; }
    double RandomDouble(const double& center, const double& radius) { return center + 2.0 * radius * RandomDouble() - radius; }

private:
    std::mt19937 engine_;
    std::uniform_int_distribution<int> dist_int_;
    std::uniform_real_distribution<float> dist_float_;
    std::uniform_real_distribution<double> dist_double_;
};

// Column-major fill in the order of RandomMatrixSequential (matrix_generator.hpp:229-248). The reference
// switches to a thread-count-dependent generator for >= 32768 elements when max_threads > 1 (it also leaves
// the remainder rows unwritten there); this build always draws the sequential stream.
template <typename T>
inline bool RandomMatrix(T* buf, const unsigned int ldim, const unsigned int height, const unsigned int width,
                         Random& rng, const T rng_center = T(0.5), const T rng_radius = T(0.5))
{
    for (unsigned int c = 0; c < width; ++c)
    {
        T* col = buf + static_cast<size_t>(c) * ldim;
        for (unsigned int r = 0; r < height; ++r) col[r] = static_cast<T>(rng.RandomDouble(rng_center, rng_radius));
    }
    return true;
}
