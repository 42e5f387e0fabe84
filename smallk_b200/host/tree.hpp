// smallk_b200 host — the HierNMF2 cluster tree.
// Same class name and public member functions as the reference's Tree<T> (hierclust/include/tree.hpp:56-158) so
// code holding a Tree<R> (the hierclust CLI, smallk::HierNmf2, FlatclustInitW in the flat step) is unchanged.
// The tree is host control flow around the GPU factorizations: node bookkeeping, document partitions by the
// sign of H(0,c) - H(1,c), topic vectors = columns of W, top terms, assignments. Storage differs from the
// reference (plain std::vector topic vectors instead of Elemental matrices).
#pragma once

#include <algorithm>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <cstdlib>
#include <new>
#include <vector>

#include "hierclust_writer.hpp"

// Row indices of the `maxterms` largest entries of v[0..height), in decreasing order of value: what the reference's
// TopTerms computes with an unstable std::sort of all `height` indices (common/include/terms.hpp:24-60). A
// partial selection gives the same list whenever the values around the cut are distinct; when they are not (ties,
// typically zeros) the answer depends on that sort's internal order, so the same std::sort call is made then.
template <typename T>
inline void TopTerms(const int maxterms, const T* v, const int height, std::vector<int>& scratch, std::vector<int>& term_indices)
{
    if (static_cast<int>(term_indices.size()) < maxterms) throw std::runtime_error("TopTerms: term array too small");
    const int keep = std::min(maxterms, height);
    scratch.resize(height);
    for (int q = 0; q < height; ++q) scratch[q] = q;
    auto by_value = [v](int a, int b) { return v[a] > v[b]; };
    bool unique = true;
    if (keep < height)
    {
        std::partial_sort(scratch.begin(), scratch.begin() + keep + 1, scratch.end(), by_value);
        for (int q = 0; q < keep && unique; ++q) unique = v[scratch[q]] > v[scratch[q + 1]];
    }
    else unique = false;
    if (!unique)
    {
        for (int q = 0; q < height; ++q) scratch[q] = q;
        std::sort(scratch.begin(), scratch.begin() + height, by_value);
    }
    for (int q = 0; q < keep; ++q) term_indices[q] = scratch[q];
}

// n zero-initialised elements straight from calloc. A topic vector of a deep node is zero on all but a few rows: pages that
// are never written stay unmapped (std::vector would write every one of them), and a later full read maps the shared zero page.
template <typename T>
class ZeroArray
{
public:
    ZeroArray() {}
    ~ZeroArray() { std::free(p_); }
    ZeroArray(const ZeroArray& o) { reset(o.n_); for (size_t i = 0; i < n_; ++i) p_[i] = o.p_[i]; }
    ZeroArray(ZeroArray&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    ZeroArray& operator=(ZeroArray o) noexcept { std::swap(p_, o.p_); std::swap(n_, o.n_); return *this; }
    void reset(const size_t n)
    {
        std::free(p_); p_ = nullptr; n_ = 0;
        if (n == 0) return;
        p_ = static_cast<T*>(std::calloc(n, sizeof(T)));
        if (!p_) throw std::bad_alloc();
        n_ = n;
    }
    T* data() { return p_; }
    const T* data() const { return p_; }
    size_t size() const { return n_; }
    T* begin() { return p_; }
    T* end() { return p_ + n_; }
    const T* begin() const { return p_; }
    const T* end() const { return p_ + n_; }
private:
    T* p_ = nullptr;
    size_t n_ = 0;
};

template <typename T>
class Tree
{
public:
    Tree() : active_nodes_(0), index0_(0), index1_(0), total_docs_(0), leaf_doc_count_(0), term_count_(0) {}

    unsigned int LeftChildIndex() { return index0_; }
    unsigned int RightChildIndex() { return index1_; }
    ZeroArray<T>& LeftChildTopicVector() { return nodes_[index0_].topic_vector; }
    ZeroArray<T>& RightChildTopicVector() { return nodes_[index1_].topic_vector; }
    std::vector<unsigned int>& LeftChildDocs() { return nodes_[index0_].docs; }
    std::vector<unsigned int>& RightChildDocs() { return nodes_[index1_].docs; }
    std::vector<unsigned int>& Outliers() { return outliers_; }
    std::vector<unsigned int>& Assignments() { return assignments_; }

    // tree.hpp:162-189
    void Init(const unsigned int num_clusters, const unsigned int node_count, const unsigned int term_count,
              const unsigned int doc_count)
    {
        (void)num_clusters;
        total_docs_ = doc_count;
        term_count_ = term_count;
        nodes_.assign(node_count, NodeRec());
        // topic vectors are allocated when a node is created (TakeTopicVectors): node_count x term_count zeros up front were
        // 320 MB of page faults at 126 nodes x 320 000 terms, all of them overwritten later
        is_leaf_.assign(node_count, false);
        active_nodes_ = 0;
        outliers_.clear(); assignments_.clear();
    }

    // tree.hpp:193-219: smallest positive and largest leaf priority, and where the largest is
    void MinMaxLeafPriorities(T& min_priority, T& max_priority, unsigned int& max_priority_index)
    {
        min_priority = std::numeric_limits<T>::max();
        max_priority = std::numeric_limits<T>::lowest();
        for (unsigned int q = 0; q < is_leaf_.size(); ++q)
        {
            if (!is_leaf_[q]) continue;
            const T p = nodes_[q].priority;
            if (p > T(0) && p < min_priority) min_priority = p;
            if (p > max_priority) { max_priority = p; max_priority_index = q; }
        }
    }

    // W: term_count x 2 column-major (ld = term_count); H: 2 x width column-major (ld = 2)
    void SplitRoot(const T* W, const T* H, const unsigned int h_width)
    {
        index0_ = 0; index1_ = 1;
        MakeLeaf(0, NONE, true); MakeLeaf(1, NONE, false);
        active_nodes_ += 2;
        for (unsigned int c = 0; c < h_width; ++c)
            nodes_[(H[2 * c] > H[2 * c + 1]) ? 0 : 1].docs.push_back(c);        // tree.hpp:254-260
        TakeTopicVectors(W);
    }

    void Split(const unsigned int node_index, const T* W, const T* H, const unsigned int h_width)
    {
        SplitDocs(node_index, H, h_width);
        TakeTopicVectors(W);
    }

    // Split with the node's W held on its own rows only (the hierclust driver's node factors): rows[r] of the full m x 2 matrix
    // is (Wc[r], Wc[count + r]); every other row is zero.
    void SplitCompact(const unsigned int node_index, const unsigned int* rows, const unsigned int count, const T* Wc, const T* H,
                      const unsigned int h_width)
    {
        SplitDocs(node_index, H, h_width);
        ZeroArray<T>& t0 = nodes_[index0_].topic_vector;
        ZeroArray<T>& t1 = nodes_[index1_].topic_vector;
        t0.reset(term_count_); t1.reset(term_count_);
        T* d0 = t0.data(); T* d1 = t1.data();
        for (unsigned int r = 0; r < count; ++r) { d0[rows[r]] = Wc[r]; d1[rows[r]] = Wc[static_cast<size_t>(count) + r]; }
    }

    // Takes back the most recent Split / SplitCompact (of node_index): the driver splits the most promising leaf ahead of time
    // while the last score is still being evaluated, and must be able to return to the tree as it was.
    void UndoSplit(const unsigned int node_index)
    {
        nodes_[index0_] = NodeRec(); nodes_[index1_] = NodeRec();
        is_leaf_[index0_] = false; is_leaf_[index1_] = false;
        active_nodes_ -= 2;
        nodes_[node_index].left_child_index = NONE; nodes_[node_index].right_child_index = NONE;
        is_leaf_[node_index] = true;
    }

    // MinMaxLeafPriorities over every leaf but skip_a and skip_b (NONE: nothing to skip); same scan, same tie rule
    void MinMaxLeafPrioritiesWithout(const unsigned int skip_a, const unsigned int skip_b, T& min_priority, T& max_priority,
                                     unsigned int& max_priority_index)
    {
        min_priority = std::numeric_limits<T>::max();
        max_priority = std::numeric_limits<T>::lowest();
        for (unsigned int q = 0; q < is_leaf_.size(); ++q)
        {
            if (!is_leaf_[q] || q == skip_a || q == skip_b) continue;
            const T p = nodes_[q].priority;
            if (p > T(0) && p < min_priority) min_priority = p;
            if (p > max_priority) { max_priority = p; max_priority_index = q; }
        }
    }

    void SetNodePriority(const unsigned int node_index, const T priority) { nodes_[node_index].priority = priority; }
    T NodePriority(const unsigned int node_index) const { return nodes_[node_index].priority; }
    bool IsLeaf(const unsigned int node_index) const { return is_leaf_[node_index]; }
    unsigned int NodeCount() const { return static_cast<unsigned int>(nodes_.size()); }

    void ComputeTopTerms(const unsigned int max_terms)
    {
        std::vector<int> scratch;
        for (auto& nd : nodes_)
        {
            if (!nd.is_valid) continue;
            nd.term_indices.resize(max_terms);
            TopTerms(static_cast<int>(max_terms), nd.topic_vector.data(), static_cast<int>(term_count_), scratch, nd.term_indices);
        }
    }

    // tree.hpp:375-410
    void ComputeAssignments()
    {
        outliers_.clear();
        assignments_.assign(total_docs_, NONE);
        leaf_doc_count_ = 0;
        for (unsigned int q = 0; q < nodes_.size(); ++q)
        {
            if (!is_leaf_[q]) continue;
            leaf_doc_count_ += static_cast<unsigned int>(nodes_[q].docs.size());
            for (unsigned int d : nodes_[q].docs) assignments_[d] = q;
        }
        for (unsigned int j = 0; j < assignments_.size(); ++j) if (NONE == assignments_[j]) outliers_.push_back(j);
    }

    // tree.hpp:414-460: leaf topic vectors, in node order, become the columns of the m x k flat-clustering W
    bool FlatclustInitW(T* Winit, const unsigned int ldim, const unsigned int m, const unsigned int k)
    {
        unsigned int leaves = 0;
        for (unsigned int q = 0; q < nodes_.size(); ++q) if (is_leaf_[q]) ++leaves;
        if (k != leaves) { std::cerr << "Insufficient number of leaf nodes for flat clustering." << std::endl; return false; }
        if (m != term_count_) { std::cerr << "Invalid W matrix height for flat clustering." << std::endl; return false; }
        unsigned int c = 0;
        for (unsigned int q = 0; q < nodes_.size(); ++q)
        {
            if (!is_leaf_[q]) continue;
            std::copy(nodes_[q].topic_vector.begin(), nodes_[q].topic_vector.end(), Winit + static_cast<size_t>(c) * ldim);
            ++c;
        }
        return k == c;
    }

    // tree.hpp:464-506: one line of node ids (-1 = outlier), a blank line, one line of outlier ids
    bool WriteAssignments(const std::string& filepath)
    {
        std::ofstream out(filepath);
        if (!out) { std::cerr << "Tree::WriteAssignments: could not open output file " << filepath << std::endl; return false; }
        out << assignments_[0];
        for (unsigned int q = 1; q < assignments_.size(); ++q)
        {
            out << ",";
            if (NONE == assignments_[q]) out << -1; else out << assignments_[q];
        }
        out << std::endl << std::endl;
        if (!outliers_.empty())
        {
            out << outliers_[0];
            for (unsigned int q = 1; q < outliers_.size(); ++q) out << ',' << outliers_[q];
            out << std::endl;
        }
        return true;
    }

    // tree.hpp:510-546
    bool WriteTree(IHierclustWriter* writer, const std::string& filepath, const std::vector<std::string>& dictionary)
    {
        std::ofstream out(filepath);
        if (!out) { std::cerr << "Tree::Write: could not open output file " << filepath << std::endl; return false; }
        writer->WriteHeader(out, leaf_doc_count_);
        for (unsigned int q = 0; q < nodes_.size(); ++q)
        {
            const NodeRec& nd = nodes_[q];
            writer->WriteNodeBegin(out, q);
            writer->WriteParentId(out, nd.parent_index);
            writer->WriteLeftChild(out, nd.is_left_child, nd.left_child_index);
            writer->WriteRightChild(out, nd.right_child_index);
            writer->WriteDocCount(out, static_cast<int>(nd.docs.size()));
            writer->WriteTopTerms(out, nd.term_indices, dictionary);
            writer->WriteNodeEnd(out);
        }
        writer->WriteFooter(out);
        return true;
    }

    void Print()
    {
        std::cout << "\n\ncluster sizes: \n\t";
        for (auto& nd : nodes_) std::cout << nd.docs.size() << "  ";
        std::cout << "\ncluster priorities: \n";
        for (auto& nd : nodes_) std::cout << nd.priority << "  ";
        std::cout << "\nleaf nodes: \n";
        int leaves = 0;
        for (unsigned int q = 0; q < is_leaf_.size(); ++q) if (is_leaf_[q]) { ++leaves; std::cout << q << ", "; }
        std::cout << "\nleaf node count: " << leaves << "\nFound " << outliers_.size() << " outliers." << std::endl;
    }

    // read-only node access for tests and bindings
    struct NodeView { int parent, left, right; bool is_left_child, is_valid, is_leaf; unsigned int doc_count; T priority; const std::vector<int>* terms; };
    NodeView Node(const unsigned int q) const
    {
        const NodeRec& nd = nodes_[q];
        return NodeView{static_cast<int>(nd.parent_index), static_cast<int>(nd.left_child_index), static_cast<int>(nd.right_child_index),
                        nd.is_left_child, nd.is_valid, static_cast<bool>(is_leaf_[q]), static_cast<unsigned int>(nd.docs.size()),
                        nd.priority, &nd.term_indices};
    }

    enum : unsigned int { NONE = 0xFFFFFFFFu };

private:
    struct NodeRec
    {
        T priority = T(0);
        unsigned int parent_index = NONE, left_child_index = NONE, right_child_index = NONE;
        bool is_valid = false, is_left_child = false;
        ZeroArray<T> topic_vector;
        std::vector<int> term_indices;
        std::vector<unsigned int> docs;
    };

    void MakeLeaf(const unsigned int q, const unsigned int parent, const bool is_left)
    {
        NodeRec& nd = nodes_[q];
        nd.parent_index = parent; nd.left_child_index = NONE; nd.right_child_index = NONE;
        nd.is_valid = true; nd.is_left_child = is_left;
        is_leaf_[q] = true;
    }
    // the two new leaves of node_index and their documents
    void SplitDocs(const unsigned int node_index, const T* H, const unsigned int h_width)
    {
        index0_ = active_nodes_; index1_ = active_nodes_ + 1;
        active_nodes_ += 2;
        nodes_[node_index].left_child_index = index0_;
        nodes_[node_index].right_child_index = index1_;
        is_leaf_[node_index] = false;
        MakeLeaf(index0_, node_index, true); MakeLeaf(index1_, node_index, false);
        const std::vector<unsigned int>& src = nodes_[node_index].docs;
        for (unsigned int c = 0; c < h_width; ++c)
            nodes_[(H[2 * c] > H[2 * c + 1]) ? index0_ : index1_].docs.push_back(src[c]);   // tree.hpp:308-314
    }
    // left child <- W(:,0), right child <- W(:,1)   (tree.hpp:332-349)
    void TakeTopicVectors(const T* W)
    {
        nodes_[index0_].topic_vector.reset(term_count_);
        nodes_[index1_].topic_vector.reset(term_count_);
        std::copy(W, W + term_count_, nodes_[index0_].topic_vector.data());
        std::copy(W + term_count_, W + 2 * static_cast<size_t>(term_count_), nodes_[index1_].topic_vector.data());
    }

    std::vector<NodeRec> nodes_;
    std::vector<bool> is_leaf_;
    unsigned int active_nodes_, index0_, index1_, total_docs_, leaf_doc_count_, term_count_;
    std::vector<unsigned int> outliers_, assignments_;
};
