// Stand-in for the part of the reference's src/Structs.h that the Demons OPERATION touches, so that
// src/Operations/RegisterImagesDemons.cc compiles unmodified in this container (the real header needs Ygor, Boost,
// Explicator, ...).  Written from the usage sites in RegisterImagesDemons.cc:39-384 and the member lists at
// src/Structs.h:654-681 (operation documentation) -- only the members used there.  NOT a product component.
#pragma once
#include <algorithm>
#include <cctype>
#include <list>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <variant>

#include "YgorImages.h"

#include "Alignment_Field.h"
#include "Alignment_Rigid.h"
#include "Alignment_TPSRPM.h"

enum class OpArgVisibility { Show, Hide };
enum class OpArgFlow { Unknown, Ingress, Egress };
enum class OpArgSamples { Examples, Exhaustive };

struct OperationArgDoc {
    std::string name, desc, default_val;
    bool expected = false;
    std::list<std::string> examples;
    std::string mimetype;
    OpArgVisibility visibility = OpArgVisibility::Show;
    OpArgFlow flow = OpArgFlow::Unknown;
    OpArgSamples samples = OpArgSamples::Examples;
};

struct OperationDoc {
    std::list<OperationArgDoc> args;
    std::string name;
    std::list<std::string> aliases;
    std::string desc;
    std::list<std::string> notes, tags;
};

// key -> value arguments of one operation invocation; keys compare case-insensitively
class OperationArgPkg {
    std::string name_;
    std::map<std::string, std::string> opts_;
    static std::string fold(std::string s) {
        std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
        return s;
    }
  public:
    explicit OperationArgPkg(std::string name = "") : name_(std::move(name)) {}
    const std::string &getName() const { return name_; }
    bool insert(const std::string &key, const std::string &val) { return opts_.emplace(fold(key), val).second; }
    void set(const std::string &key, const std::string &val) { opts_[fold(key)] = val; }
    std::optional<std::string> getValueStr(const std::string &key) const {
        const auto it = opts_.find(fold(key));
        if (it == opts_.end()) return std::nullopt;
        return it->second;
    }
};

struct Image_Array {
    planar_image_collection<float, double> imagecoll;
    std::string filename;
};


struct Transform3 {
    std::variant<std::monostate, affine_transform<double>, thin_plate_spline, deformation_field> transform;
    std::map<std::string, std::string> metadata;
};

struct Contour_Data {
    std::list<contour_collection<double>> ccs;
};

class Drover {
  public:
    std::shared_ptr<Contour_Data> contour_data;
    std::list<std::shared_ptr<Image_Array>> image_data;
    std::list<std::shared_ptr<Transform3>> trans_data;
};
