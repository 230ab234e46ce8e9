//---------------------------------------------------------------------------//
// Native ORANGE construction: .org.json -> the geometry columns of a problem image.
//
// SURVEY 8(f)2. The reference builds its runtime geometry in OrangeParams(OrangeInput&&)
// (src/orange/OrangeParams.cc:137-209) through UnitInserter, RectArrayInserter,
// UniverseInserter, SurfacesRecordBuilder, TransformRecordInserter and BIHBuilder /
// BIHPartitioner (src/orange/detail/*.cc); until now every .b2img here was produced by that
// code through the oracle harness. This file does the same construction without the
// reference: it reads the reference's ORANGE JSON input format
// (src/orange/OrangeInputIO.json.cc) and writes the `geo.*` columns that
// celeritas_b200/host/CoreParams.cc loads, so `celer-sim-b200` and
// `b200_params_create_from_org_json` can open a geometry file directly.
//
// The result is the reference's, column for column (tests/test_cpu_orange_builder.py
// compares every geo.* column of every geometry under data/geometry with the image the
// reference built): same record order, the same de-duplication of repeated ranges
// (DedupeCollectionBuilder, corecel/data/DedupeCollectionBuilder.hh: every builder keeps its
// own set of inserted ranges), the same bumped float bounding boxes, and the same BIH tree
// (partition candidates, cost function in float, node arrangement).
//---------------------------------------------------------------------------//
#include "OrangeBuilder.hh"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <variant>
#include <nlohmann/json.hpp>

namespace celeritas_b200
{
namespace
{
using json = nlohmann::json;
using b200::Image;
constexpr uint32_t INVALID_ID = 0xffffffffu;

// Logic tokens (orange/OrangeTypes.hh:244-254, logic_int = uint32: lbegin = ~6); the image
// keeps the reference's 32-bit tokens, the loader narrows them for the device
constexpr uint32_t LBEGIN = 0xfffffff9u, LTRUE = 0xfffffffbu, LOR = 0xfffffffcu,
                   LAND = 0xfffffffdu, LNOT = 0xfffffffeu;
// VolumeRecord::Flags (orange/OrangeData.hh)
constexpr uint32_t F_INTERNAL = 1, F_IMPLICIT = 2, F_SIMPLE_SAFETY = 4, F_EMBEDDED = 8;

[[noreturn]] void fail(std::string const& what)
{
    throw std::runtime_error("ORANGE input: " + what);
}

//! Append ranges to a column, returning an earlier identical range if this builder has
//! inserted one (DedupeCollectionBuilder::insert_back)
template<class T>
class Dedupe
{
  public:
    explicit Dedupe(std::vector<T>* column) : col_(column) {}
    uint32_t insert(std::vector<T> const& values)
    {
        std::string key(reinterpret_cast<char const*>(values.data()), values.size() * sizeof(T));
        auto found = seen_.find(key);
        if (found != seen_.end())
            return found->second;
        uint32_t const start = col_->size();
        col_->insert(col_->end(), values.begin(), values.end());
        seen_.emplace(std::move(key), start);
        return start;
    }

  private:
    std::vector<T>* col_;
    std::map<std::string, uint32_t> seen_;
};

char const* const surface_names[] = {"px", "py", "pz", "cxc", "cyc", "czc", "sc", "cx", "cy",
                                     "cz", "p", "s", "kx", "ky", "kz", "sq", "gq", "inv"};
constexpr int surface_sizes[] = {1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 4, 4, 4, 4, 4, 7, 10, 0};

uint8_t surface_type(std::string const& name)
{
    for (uint8_t t = 0; t < 18; ++t)
        if (name == surface_names[t])
            return t;
    fail("unknown surface type '" + name + "'");
}
//! Intersections per surface type (surf/*.hh: Intersections array size)
uint32_t num_intersections(uint8_t t)
{
    return (t <= 2 || t == 10) ? 1u : 2u;
}
//! Surface types whose safety distance is exact (surf/*.hh: simple_safety())
bool simple_safety(uint8_t t)
{
    return t <= 6 || t == 10 || t == 11;
}

struct FBox  // BoundingBox<float>
{
    std::array<float, 3> lo, hi;
    static FBox null()
    {
        float const inf = std::numeric_limits<float>::infinity();
        return {{inf, inf, inf}, {-inf, -inf, -inf}};
    }
    static FBox infinite()
    {
        float const inf = std::numeric_limits<float>::infinity();
        return {{-inf, -inf, -inf}, {inf, inf, inf}};
    }
    bool is_infinite() const
    {
        float const inf = std::numeric_limits<float>::infinity();
        for (int a = 0; a < 3; ++a)
            if (lo[a] != -inf || hi[a] != inf)
                return false;
        return true;
    }
};

FBox box_union(FBox const& a, FBox const& b)
{
    FBox r;
    for (int ax = 0; ax < 3; ++ax)
    {
        r.lo[ax] = std::min(a.lo[ax], b.lo[ax]);
        r.hi[ax] = std::max(a.hi[ax], b.hi[ax]);
    }
    return r;
}

float surface_area(FBox const& b)
{
    float const lx = b.hi[0] - b.lo[0], ly = b.hi[1] - b.lo[1], lz = b.hi[2] - b.lo[2];
    return 2 * (lx * ly + lx * lz + ly * lz);
}

//! Volume bounding box of the input (double), null when absent from a background volume
struct DBox
{
    bool valid{false};
    std::array<double, 3> lo{}, hi{};
};

DBox read_bbox(json const& j, bool absent_is_infinite)
{
    double const inf = std::numeric_limits<double>::infinity();
    DBox b;
    auto it = j.find("bbox");
    if (it == j.end())
    {
        if (absent_is_infinite)
        {
            b.valid = true;
            b.lo = {-inf, -inf, -inf};
            b.hi = {inf, inf, inf};
        }
        return b;
    }
    if (it->is_null())
        return b;
    if (!it->is_array() || it->size() != 2)
        fail("bounding box must have lower and upper extents");
    b.lo = (*it)[0].get<std::array<double, 3>>();
    b.hi = (*it)[1].get<std::array<double, 3>>();
    double const big = std::numeric_limits<double>::max();
    for (int ax = 0; ax < 3; ++ax)
    {
        if (std::fabs(b.lo[ax]) == big)
            b.lo[ax] = std::copysign(inf, b.lo[ax]);
        if (std::fabs(b.hi[ax]) == big)
            b.hi[ax] = std::copysign(inf, b.hi[ax]);
    }
    b.valid = b.lo[0] <= b.hi[0] && b.lo[1] <= b.hi[1] && b.lo[2] <= b.hi[2];
    return b;
}

//! BoundingBoxBumper<float, double> with twice the geometry tolerance (make_bumper)
struct Bumper
{
    double rel, abs;
    float bump(double value, int sign) const
    {
        double const bumped = value + sign * std::max(abs, rel * std::fabs(value));
        return std::nextafter(static_cast<float>(bumped),
                              sign * std::numeric_limits<float>::infinity());
    }
    FBox operator()(DBox const& b) const
    {
        FBox r;
        for (int ax = 0; ax < 3; ++ax)
        {
            r.lo[ax] = this->bump(b.lo[ax], -1);
            r.hi[ax] = this->bump(b.hi[ax], +1);
        }
        return r;
    }
};

std::vector<uint32_t> parse_logic(std::string const& s)
{
    std::vector<uint32_t> result;
    uint32_t id = 0;
    bool reading = false;
    for (char v : s)
    {
        if (v >= '0' && v <= '9')
        {
            if (!reading)
            {
                id = 0;
                reading = true;
            }
            id = 10 * id + (v - '0');
            continue;
        }
        if (reading)
        {
            result.push_back(id);
            reading = false;
        }
        switch (v)
        {
            case '*': result.push_back(LTRUE); continue;
            case '|': result.push_back(LOR); continue;
            case '&': result.push_back(LAND); continue;
            case '~': result.push_back(LNOT); continue;
            case ' ': continue;
            default: fail(std::string("unexpected token '") + v + "' while parsing logic string");
        }
    }
    if (reading)
        result.push_back(id);
    return result;
}

int logic_depth(std::vector<uint32_t> const& logic)
{
    int max_depth = 1, cur = 0;
    for (uint32_t id : logic)
    {
        bool const is_operator = id >= LBEGIN;
        if (!is_operator || id == LTRUE)
            ++cur;
        else if (id == LAND || id == LOR)
        {
            max_depth = std::max(cur, max_depth);
            --cur;
        }
    }
    return cur == 1 ? max_depth : -1;
}

//! The name part of "name@ext" (corecel/io/Label.cc:40-51)
std::string label_name(std::string const& text)
{
    auto pos = text.rfind('@');
    return pos == std::string::npos ? text : text.substr(0, pos);
}

//! Label written to the image: the name; volumes of one unit that share a name keep their
//! explicit extension ("box@1" .. "box@4") so that detectors can be attached to one of them
std::string image_label(std::vector<std::string> const& labels, size_t i)
{
    if (labels.empty())
        return {};
    std::string const name = label_name(labels[i]);
    if (name.size() + 1 >= labels[i].size())
        return name;  // no extension
    size_t same = 0;
    for (std::string const& other : labels)
        same += label_name(other) == name;
    return same > 1 ? labels[i] : name;
}

//! A daughter placement: universe + transform data (0, 3 or 12 reals)
struct DaughterInput
{
    uint32_t universe{INVALID_ID};
    std::vector<double> transform;
};

std::vector<double> translation_or_none(double const* t)
{
    if (t[0] == 0 && t[1] == 0 && t[2] == 0)
        return {};
    return {t[0], t[1], t[2]};
}

//---------------------------------------------------------------------------//
//! All columns under construction (HostVal<OrangeParamsData>)
struct Columns
{
    uint32_t max_depth{0}, max_faces{1}, max_intersections{1}, max_logic_depth{0};
    double tol_rel{0}, tol_abs{0};
    std::vector<uint8_t> universe_type;
    std::vector<uint32_t> universe_index, surface_offset{0}, volume_offset{0};
    std::vector<uint32_t> local_surface_ids, local_volume_ids, real_ids, logic_ints;
    std::vector<double> reals;
    std::vector<uint8_t> surface_types;
    std::vector<uint32_t> vol_face_begin, vol_face_end, vol_logic_begin, vol_logic_end,
        vol_max_isect, vol_flags, vol_daughter;
    std::vector<uint32_t> conn_begin, conn_end;
    std::vector<uint32_t> daughter_universe, daughter_transform;
    std::vector<uint8_t> transform_type;
    std::vector<uint32_t> transform_offset;
    std::vector<float> bih_bboxes;
    std::vector<uint32_t> bih_local_volume_ids;
    std::vector<uint32_t> inner_parent, inner_axis, inner_left_child, inner_right_child;
    std::vector<float> inner_left_pos, inner_right_pos;
    std::vector<uint32_t> leaf_parent, leaf_begin, leaf_end;
    std::vector<uint32_t> simple_units, rect_arrays;
    std::string volume_labels;
    uint32_t num_simple{0}, num_rect{0};
};

//! TransformRecordInserter: one per universe inserter, each with its own de-duplication
class TransformInserter
{
  public:
    explicit TransformInserter(Columns* c) : c_(c), reals_(&c->reals) {}
    uint32_t operator()(std::vector<double> const& data)
    {
        if (data.empty())
        {
            if (null_ != INVALID_ID)
                return null_;
            null_ = c_->transform_type.size();
        }
        uint8_t const type = data.empty() ? 0 : (data.size() == 3 ? 1 : 2);
        uint32_t const offset = reals_.insert(data);
        c_->transform_type.push_back(type);
        c_->transform_offset.push_back(offset);
        return c_->transform_type.size() - 1;
    }

  private:
    Columns* c_;
    Dedupe<double> reals_;
    uint32_t null_{INVALID_ID};
};

//---------------------------------------------------------------------------//
// BIH (detail/BIHBuilder.cc, detail/BIHPartitioner.cc)
//---------------------------------------------------------------------------//
struct BihNode
{
    bool leaf{true};
    uint32_t parent{INVALID_ID};
    // inner
    uint32_t axis{0};
    float left_pos{0}, right_pos{0};
    uint32_t left_child{INVALID_ID}, right_child{INVALID_ID};
    // leaf
    uint32_t vol_begin{0}, vol_end{0};
};

struct Partition
{
    bool valid{false};
    uint32_t axis{0};
    std::vector<uint32_t> indices[2];
    FBox boxes[2];
};

class BihBuilder
{
  public:
    explicit BihBuilder(Columns* c) : c_(c) {}

    //! Returns {bbox begin, inner begin, inner count, leaf begin, leaf count, inf begin, count}
    std::array<uint32_t, 7> operator()(std::vector<FBox> boxes)
    {
        boxes_ = std::move(boxes);
        centers_.resize(boxes_.size());
        for (size_t i = 0; i < boxes_.size(); ++i)
            for (int ax = 0; ax < 3; ++ax)
                centers_[i][ax] = (boxes_[i].lo[ax] + boxes_[i].hi[ax]) / 2;
        std::vector<uint32_t> finite, infinite;
        for (uint32_t i = 0; i < boxes_.size(); ++i)
            (boxes_[i].is_infinite() ? infinite : finite).push_back(i);

        std::array<uint32_t, 7> out{};
        out[0] = c_->bih_bboxes.size() / 6;
        for (FBox const& b : boxes_)
        {
            c_->bih_bboxes.insert(c_->bih_bboxes.end(), b.lo.begin(), b.lo.end());
            c_->bih_bboxes.insert(c_->bih_bboxes.end(), b.hi.begin(), b.hi.end());
        }
        out[5] = c_->bih_local_volume_ids.size();
        out[6] = infinite.size();
        c_->bih_local_volume_ids.insert(
            c_->bih_local_volume_ids.end(), infinite.begin(), infinite.end());

        std::vector<BihNode> nodes;
        if (!finite.empty())
            this->construct(finite, &nodes, INVALID_ID);
        else
            nodes.push_back(BihNode{});  // one empty leaf
        // arrange: inner nodes first, then leaves, ids remapped
        std::vector<uint32_t> new_index(nodes.size());
        uint32_t ninner = 0, nleaf = 0;
        for (size_t i = 0; i < nodes.size(); ++i)
            new_index[i] = nodes[i].leaf ? nleaf++ : ninner++;
        for (size_t i = 0; i < nodes.size(); ++i)
            if (nodes[i].leaf)
                new_index[i] += ninner;
        auto remap = [&](uint32_t id) { return id == INVALID_ID ? id : new_index[id]; };
        out[1] = c_->inner_parent.size();
        out[2] = ninner;
        out[3] = c_->leaf_parent.size();
        out[4] = nleaf;
        for (BihNode const& n : nodes)
        {
            if (n.leaf)
                continue;
            c_->inner_parent.push_back(remap(n.parent));
            c_->inner_axis.push_back(n.axis);
            c_->inner_left_pos.push_back(n.left_pos);
            c_->inner_left_child.push_back(remap(n.left_child));
            c_->inner_right_pos.push_back(n.right_pos);
            c_->inner_right_child.push_back(remap(n.right_child));
        }
        for (BihNode const& n : nodes)
        {
            if (!n.leaf)
                continue;
            c_->leaf_parent.push_back(remap(n.parent));
            c_->leaf_begin.push_back(n.vol_begin);
            c_->leaf_end.push_back(n.vol_end);
        }
        return out;
    }

  private:
    Columns* c_;
    std::vector<FBox> boxes_;
    std::vector<std::array<float, 3>> centers_;

    void construct(std::vector<uint32_t> const& indices, std::vector<BihNode>* nodes, uint32_t parent)
    {
        size_t const current = nodes->size();
        nodes->resize(current + 1);
        Partition p = this->partition(indices);
        if (p.valid)
        {
            BihNode node;
            node.leaf = false;
            node.parent = parent;
            node.axis = p.axis;
            node.left_pos = p.boxes[0].hi[p.axis];
            node.right_pos = p.boxes[1].lo[p.axis];
            node.left_child = nodes->size();
            this->construct(p.indices[0], nodes, current);
            node.right_child = nodes->size();
            this->construct(p.indices[1], nodes, current);
            (*nodes)[current] = node;
        }
        else
        {
            BihNode node;
            node.parent = parent;
            node.vol_begin = c_->bih_local_volume_ids.size();
            c_->bih_local_volume_ids.insert(
                c_->bih_local_volume_ids.end(), indices.begin(), indices.end());
            node.vol_end = c_->bih_local_volume_ids.size();
            (*nodes)[current] = node;
        }
    }

    Partition partition(std::vector<uint32_t> const& indices) const
    {
        Partition best;
        double best_cost = std::numeric_limits<double>::infinity();
        constexpr uint32_t candidates_per_axis = 3;
        for (uint32_t ax = 0; ax < 3; ++ax)
        {
            // sorted centres, soft-unique (SoftEqual<double>: rel 1e-12, abs 1e-14)
            std::vector<double> centers;
            for (uint32_t id : indices)
                centers.push_back(centers_[id][ax]);
            std::sort(centers.begin(), centers.end());
            auto soft_equal = [](double a, double b) {
                double const rel = 1.0e-12 * std::fmax(std::fabs(a), std::fabs(b));
                return std::fabs(a - b) < std::fmax(1.0e-14, rel);
            };
            centers.erase(std::unique(centers.begin(), centers.end(), soft_equal), centers.end());
            uint32_t const step = std::max<uint32_t>(centers.size() / (candidates_per_axis + 1), 1u);
            for (uint32_t i = step; i < centers.size(); i += step)
            {
                double const position = (centers[i - 1] + centers[i]) / 2;
                Partition p;
                p.valid = true;
                p.axis = ax;
                for (uint32_t id : indices)
                    p.indices[centers_[id][ax] < position ? 0 : 1].push_back(id);
                for (int e = 0; e < 2; ++e)
                {
                    p.boxes[e] = FBox::null();
                    for (uint32_t id : p.indices[e])
                        p.boxes[e] = box_union(p.boxes[e], boxes_[id]);
                }
                // cost in float, as the reference: area * count + area * count
                double const cost = surface_area(p.boxes[0]) * p.indices[0].size()
                                    + surface_area(p.boxes[1]) * p.indices[1].size();
                if (cost < best_cost)
                {
                    best = std::move(p);
                    best_cost = cost;
                }
            }
        }
        return best;
    }
};

//---------------------------------------------------------------------------//
class Builder
{
  public:
    Builder()
        : surface_reals_(&c_.reals)
        , unit_transforms_(&c_)
        , rect_transforms_(&c_)
        , faces_(&c_.local_surface_ids)
        , neighbours_(&c_.local_volume_ids)
        , logic_(&c_.logic_ints)
        , grid_reals_(&c_.reals)
        , bih_(&c_)
    {
    }

    Image operator()(json const& j)
    {
        auto fmt = j.value("_format", std::string());
        if (fmt != "orange" && fmt != "ORANGE" && fmt != "SCALE ORANGE")
            fail("unknown format '" + fmt + "'");
        if (auto it = j.find("tol"); it != j.end())
        {
            c_.tol_rel = it->at("rel").get<double>();
            c_.tol_abs = it->at("abs").get<double>();
            if (!(c_.tol_rel > 0 && c_.tol_rel < 1) || !(c_.tol_abs > 0))
                fail("tolerance is out of range");
        }
        else
        {
            c_.tol_rel = 1.5e-8;  // Tolerance<>::from_default(): sqrt(epsilon), unit length
            c_.tol_abs = 1.5e-8;
        }
        bumper_ = Bumper{2 * c_.tol_rel, 2 * c_.tol_abs};
        json const& universes = j.at("universes");
        if (universes.empty())
            fail("no universes");
        // max_depth (detail/DepthCalculator.cc)
        std::map<uint32_t, uint32_t> depths;
        c_.max_depth = this->depth(universes, 0, &depths);
        for (json const& u : universes)
        {
            std::string const type = u.at("_type").get<std::string>();
            if (type == "unit" || type == "simple unit")
                this->insert_unit(u);
            else if (type == "rectarray" || type == "rectangular array")
                this->insert_rect_array(u);
            else
                fail("unsupported universe type '" + type + "'");
        }
        if (universes[0].at("_type").get<std::string>().find("unit") == std::string::npos)
            fail("global universe is not a SimpleUnit");
        return this->finish();
    }

  private:
    Columns c_;
    Dedupe<double> surface_reals_;
    TransformInserter unit_transforms_, rect_transforms_;
    Dedupe<uint32_t> faces_, neighbours_, logic_;
    Dedupe<double> grid_reals_;
    BihBuilder bih_;
    Bumper bumper_{};

    static std::vector<DaughterInput> rect_daughters(json const& u)
    {
        if (u.contains("transforms"))
            fail("rect arrays with 'transforms' are not supported by the input format");
        std::vector<uint32_t> parents;
        if (auto it = u.find("parent_cells"); it != u.end())
            parents = it->get<std::vector<uint32_t>>();
        auto daughters = u.at("daughters").get<std::vector<uint32_t>>();
        auto translations = u.at("translations").get<std::vector<double>>();
        if (3 * daughters.size() != translations.size())
            fail("field 'translations' is not 3x length of 'daughters'");
        std::vector<DaughterInput> result(daughters.size());
        for (size_t i = 0; i < daughters.size(); ++i)
        {
            size_t const parent = parents.empty() ? i : parents[i];
            result.at(parent) = {daughters[i], translation_or_none(&translations[3 * i])};
        }
        return result;
    }

    static std::map<uint32_t, DaughterInput> unit_daughters(json const& u)
    {
        std::map<uint32_t, DaughterInput> result;
        for (char const* key : {"parent_volumes", "parent_cells"})
        {
            auto it = u.find(key);
            if (it == u.end())
                continue;
            auto parents = it->get<std::vector<uint32_t>>();
            auto daughters = u.at("daughters").get<std::vector<uint32_t>>();
            if (parents.size() != daughters.size())
                fail(std::string("fields '") + key + "' and 'daughters' have different lengths");
            std::vector<std::vector<double>> transforms;
            if (auto t = u.find("transforms"); t != u.end())
            {
                for (json const& e : *t)
                {
                    auto data = e.get<std::vector<double>>();
                    if (data.size() != 0 && data.size() != 3 && data.size() != 12)
                        fail("invalid number of elements in transform");
                    transforms.push_back(std::move(data));
                }
            }
            else if (auto tr = u.find("translations"); tr != u.end())
            {
                auto flat = tr->get<std::vector<double>>();
                if (flat.size() != 3 * parents.size())
                    fail("field 'translations' is not 3x length of the parents");
                for (size_t i = 0; i < parents.size(); ++i)
                    transforms.push_back(translation_or_none(&flat[3 * i]));
            }
            else
            {
                fail("missing 'transforms' or 'translations'");
            }
            if (transforms.size() != parents.size())
                fail("one transform per daughter is required");
            for (size_t i = 0; i < parents.size(); ++i)
                result.emplace(parents[i], DaughterInput{daughters[i], transforms[i]});
        }
        return result;
    }

    uint32_t depth(json const& universes, uint32_t uid, std::map<uint32_t, uint32_t>* memo)
    {
        if (uid >= universes.size())
            fail("daughter universe id out of range");
        auto found = memo->find(uid);
        if (found != memo->end())
            return found->second;
        (*memo)[uid] = 0;
        json const& u = universes[uid];
        uint32_t deepest = 0;
        std::string const type = u.at("_type").get<std::string>();
        if (type == "unit" || type == "simple unit")
        {
            for (auto const& kv : unit_daughters(u))
                deepest = std::max(deepest, this->depth(universes, kv.second.universe, memo));
        }
        else
        {
            for (DaughterInput const& d : rect_daughters(u))
                deepest = std::max(deepest, this->depth(universes, d.universe, memo));
        }
        (*memo)[uid] = deepest + 1;
        return deepest + 1;
    }

    void register_universe(uint8_t type, uint32_t num_surfaces, uint32_t num_volumes)
    {
        c_.universe_type.push_back(type);
        c_.universe_index.push_back(type == 0 ? c_.num_simple++ : c_.num_rect++);
        c_.surface_offset.push_back(c_.surface_offset.back() + num_surfaces);
        c_.volume_offset.push_back(c_.volume_offset.back() + num_volumes);
    }

    void insert_unit(json const& u)
    {
        std::string const unit_name = u.at("md").at("name").get<std::string>();
        // surfaces (SurfacesRecordBuilder)
        json const& surf = u.at("surfaces");
        auto types = surf.at("types").get<std::vector<std::string>>();
        auto data = surf.at("data").get<std::vector<double>>();
        std::vector<uint32_t> sizes;
        if (auto it = surf.find("sizes"); it != surf.end())
            sizes = it->get<std::vector<uint32_t>>();
        uint32_t const surf_begin = c_.surface_types.size();
        uint32_t const real_id_begin = c_.real_ids.size();
        std::vector<uint8_t> local_types;
        size_t at = 0;
        for (size_t i = 0; i < types.size(); ++i)
        {
            uint8_t const t = surface_type(types[i]);
            if (t == 17)
                fail("runtime involute support");
            size_t const n = surface_sizes[t];
            if (!sizes.empty() && sizes[i] != n)
                fail("surface '" + types[i] + "' has the wrong number of reals");
            if (at + n > data.size())
                fail("surface data is too short");
            std::vector<double> reals(data.begin() + at, data.begin() + at + n);
            at += n;
            c_.surface_types.push_back(t);
            c_.real_ids.push_back(surface_reals_.insert(reals));
            local_types.push_back(t);
        }
        if (at != data.size())
            fail("surface data is too long");

        // volumes (UnitInserter::insert_volume)
        json const* volumes = nullptr;
        for (char const* key : {"volumes", "cells"})
            if (auto it = u.find(key); it != u.end())
            {
                volumes = &*it;
                break;
            }
        if (!volumes || volumes->empty())
            fail("unit '" + unit_name + "' has no volumes");
        std::vector<std::string> labels;
        for (char const* key : {"volume_labels", "cell_names"})
            if (auto it = u.find(key); it != u.end())
            {
                labels = it->get<std::vector<std::string>>();
                break;
            }
        if (!labels.empty() && labels.size() != volumes->size())
            fail("incorrect size for volume labels");
        auto daughters = unit_daughters(u);
        uint32_t const vol_begin = c_.vol_flags.size();
        std::vector<std::set<uint32_t>> connectivity(types.size());
        std::vector<FBox> boxes;
        bool unit_simple_safety = true;
        bool last_is_background = false;
        for (uint32_t i = 0; i < volumes->size(); ++i)
        {
            json const& v = (*volumes)[i];
            auto faces = v.at("faces").get<std::vector<uint32_t>>();
            if (!std::is_sorted(faces.begin(), faces.end())
                || (!faces.empty() && faces.back() >= types.size()))
                fail("volume faces must be sorted local surface ids");
            uint32_t flags = v.value("flags", 0u);
            bool background = false;
            if (auto it = v.find("zorder"); it != v.end())
            {
                // a letter, or (backward compatibility) the ZOrder enum value: 1 = background
                if (it->is_string())
                    background = (it->get<std::string>() == "B");
                else
                    background = (it->get<int64_t>() == 1);
            }
            std::vector<uint32_t> logic;
            DBox bbox;
            if (background)
            {
                logic = {LTRUE, LNOT};
            }
            else
            {
                logic = parse_logic(v.at("logic").get<std::string>());
                bbox = read_bbox(v, true);
            }
            bool simple = true;
            uint32_t max_isect = 0;
            for (uint32_t f : faces)
            {
                simple = simple && simple_safety(local_types[f]);
                max_isect += num_intersections(local_types[f]);
            }
            if (v.contains("obz"))
                fail("oriented bounding zones are not supported");
            uint32_t const fb = faces_.insert(faces);
            uint32_t const lb = logic_.insert(logic);
            if (simple)
                flags |= F_SIMPLE_SAFETY;
            int const depth = logic_depth(logic);
            if (depth <= 0)
                fail("invalid logic definition: operators do not balance");
            c_.max_faces = std::max<uint32_t>(c_.max_faces, faces.size());
            c_.max_intersections = std::max(c_.max_intersections, max_isect);
            c_.max_logic_depth = std::max<uint32_t>(c_.max_logic_depth, depth);

            boxes.push_back(bbox.valid ? bumper_(bbox) : FBox::infinite());
            uint32_t daughter = INVALID_ID;
            if (auto d = daughters.find(i); d != daughters.end())
            {
                daughter = c_.daughter_universe.size();
                uint32_t const transform = unit_transforms_(d->second.transform);
                c_.daughter_universe.push_back(d->second.universe);
                c_.daughter_transform.push_back(transform);
                flags |= F_EMBEDDED;
            }
            if (!(flags & F_IMPLICIT))
                for (uint32_t f : faces)
                    connectivity[f].insert(i);
            c_.vol_face_begin.push_back(fb);
            c_.vol_face_end.push_back(fb + faces.size());
            c_.vol_logic_begin.push_back(lb);
            c_.vol_logic_end.push_back(lb + logic.size());
            c_.vol_max_isect.push_back(max_isect);
            c_.vol_flags.push_back(flags);
            c_.vol_daughter.push_back(daughter);
            // supports_simple_safety (detail/UnitInserter.cc:84-89), over all but the exterior
            if (i > 0)
                unit_simple_safety
                    = unit_simple_safety
                      && ((flags & F_IMPLICIT)
                          || ((flags & F_SIMPLE_SAFETY) && !(flags & F_INTERNAL)));
            last_is_background = background;
            c_.volume_labels += image_label(labels, i) + "\n";
        }
        auto const tree = bih_(std::move(boxes));
        uint32_t const conn_begin = c_.conn_begin.size();
        for (auto const& neighbours : connectivity)
        {
            std::vector<uint32_t> ids(neighbours.begin(), neighbours.end());
            uint32_t const b = neighbours_.insert(ids);
            c_.conn_begin.push_back(b);
            c_.conn_end.push_back(b + ids.size());
        }
        uint32_t const row[16] = {surf_begin,
                                  uint32_t(c_.surface_types.size()),
                                  real_id_begin,
                                  conn_begin,
                                  vol_begin,
                                  uint32_t(volumes->size()),
                                  last_is_background ? uint32_t(volumes->size() - 1) : INVALID_ID,
                                  unit_simple_safety ? 1u : 0u,
                                  tree[0],
                                  tree[1],
                                  tree[2],
                                  tree[3],
                                  tree[4],
                                  tree[5],
                                  tree[6],
                                  0u};
        c_.simple_units.insert(c_.simple_units.end(), row, row + 16);
        this->register_universe(0, types.size(), volumes->size());
    }

    void insert_rect_array(json const& u)
    {
        double const inf = std::numeric_limits<double>::infinity();
        uint32_t dims[3], grid_begin[3], grid_end[3], sizes[3];
        uint32_t num_volumes = 1, num_surfaces = 0;
        char const* const axes[] = {"x", "y", "z"};
        for (int ax = 0; ax < 3; ++ax)
        {
            auto grid = u.at(axes[ax]).get<std::vector<double>>();
            if (grid.size() < 2)
                fail(std::string("grid for ") + axes[ax] + " axis is too small");
            if (!std::is_sorted(grid.begin(), grid.end()))
                fail(std::string("grid for ") + axes[ax] + " axis is not monotonically increasing");
            grid.front() = -inf;
            grid.back() = inf;
            sizes[ax] = grid.size();
            dims[ax] = grid.size() - 1;
            num_volumes *= dims[ax];
            num_surfaces += grid.size();
            grid_begin[ax] = grid_reals_.insert(grid);
            grid_end[ax] = grid_begin[ax] + grid.size();
        }
        auto daughters = rect_daughters(u);
        if (daughters.size() != num_volumes)
            fail("number of input daughters does not match number of volumes");
        uint32_t const daughter_begin = c_.daughter_universe.size();
        // transforms first, then the contiguous daughter records
        std::vector<uint32_t> transforms;
        for (DaughterInput const& d : daughters)
            transforms.push_back(rect_transforms_(d.transform));
        for (size_t i = 0; i < daughters.size(); ++i)
        {
            c_.daughter_universe.push_back(daughters[i].universe);
            c_.daughter_transform.push_back(transforms[i]);
        }
        uint32_t const row[16] = {daughter_begin,
                                  num_volumes,
                                  dims[0],
                                  dims[1],
                                  dims[2],
                                  grid_begin[0],
                                  grid_end[0],
                                  grid_begin[1],
                                  grid_end[1],
                                  grid_begin[2],
                                  grid_end[2],
                                  0u,
                                  sizes[0],
                                  sizes[0] + sizes[1],
                                  sizes[0] + sizes[1] + sizes[2],
                                  0u};
        c_.rect_arrays.insert(c_.rect_arrays.end(), row, row + 16);
        for (uint32_t i = 0; i < dims[0]; ++i)
            for (uint32_t j = 0; j < dims[1]; ++j)
                for (uint32_t k = 0; k < dims[2]; ++k)
                    c_.volume_labels += "{" + std::to_string(i) + "," + std::to_string(j) + ","
                                        + std::to_string(k) + "}\n";
        this->register_universe(1, num_surfaces, num_volumes);
    }

    Image finish()
    {
        Image img;
        img.put("geo.scalars",
                std::vector<uint32_t>{c_.max_depth, c_.max_faces, c_.max_intersections,
                                      c_.max_logic_depth});
        img.put("geo.tol", std::vector<double>{c_.tol_rel, c_.tol_abs});
        img.put("geo.universe_type", c_.universe_type);
        img.put("geo.universe_index", c_.universe_index);
        img.put("geo.universe_surface_offset", c_.surface_offset);
        img.put("geo.universe_volume_offset", c_.volume_offset);
        img.put("geo.local_surface_ids", c_.local_surface_ids);
        img.put("geo.local_volume_ids", c_.local_volume_ids);
        img.put("geo.real_ids", c_.real_ids);
        img.put("geo.logic_ints", c_.logic_ints);
        img.put("geo.reals", c_.reals);
        img.put("geo.surface_types", c_.surface_types);
        img.put("geo.vol_face_begin", c_.vol_face_begin);
        img.put("geo.vol_face_end", c_.vol_face_end);
        img.put("geo.vol_logic_begin", c_.vol_logic_begin);
        img.put("geo.vol_logic_end", c_.vol_logic_end);
        img.put("geo.vol_max_isect", c_.vol_max_isect);
        img.put("geo.vol_flags", c_.vol_flags);
        img.put("geo.vol_daughter", c_.vol_daughter);
        img.put("geo.conn_begin", c_.conn_begin);
        img.put("geo.conn_end", c_.conn_end);
        img.put("geo.daughter_universe", c_.daughter_universe);
        img.put("geo.daughter_transform", c_.daughter_transform);
        img.put("geo.transform_type", c_.transform_type);
        img.put("geo.transform_offset", c_.transform_offset);
        img.put("geo.bih_bboxes", c_.bih_bboxes);
        img.put("geo.bih_local_volume_ids", c_.bih_local_volume_ids);
        img.put("geo.bih_inner_parent", c_.inner_parent);
        img.put("geo.bih_inner_axis", c_.inner_axis);
        img.put("geo.bih_inner_left_pos", c_.inner_left_pos);
        img.put("geo.bih_inner_left_child", c_.inner_left_child);
        img.put("geo.bih_inner_right_pos", c_.inner_right_pos);
        img.put("geo.bih_inner_right_child", c_.inner_right_child);
        img.put("geo.bih_leaf_parent", c_.leaf_parent);
        img.put("geo.bih_leaf_vol_begin", c_.leaf_begin);
        img.put("geo.bih_leaf_vol_end", c_.leaf_end);
        img.put("geo.simple_units", c_.simple_units);
        img.put("geo.rect_arrays", c_.rect_arrays);
        img.put_string("geo.volume_labels", c_.volume_labels);
        return img;
    }
};
}  // namespace

b200::Image build_orange_image(std::string const& org_json_path)
{
    std::ifstream in(org_json_path);
    if (!in)
        throw std::runtime_error("cannot open ORANGE input '" + org_json_path + "'");
    json j;
    try
    {
        in >> j;
    }
    catch (json::exception const& e)
    {
        throw std::runtime_error("ORANGE input '" + org_json_path + "' is not JSON: " + e.what());
    }
    Image img;
    try
    {
        img = Builder{}(j);
    }
    catch (json::exception const& e)
    {
        fail(std::string(e.what()) + " in '" + org_json_path + "'");
    }
    img.put_string("config",
                   json{{"problem", "geometry"}, {"geometry_file", org_json_path},
                        {"built_by", "celeritas_b200 OrangeBuilder"}}
                       .dump());
    return img;
}
}  // namespace celeritas_b200
