// tiny_obj_loader.h — adapter exposing the tinyobj API subset Scene::loadObj uses (scene.cpp:52-71,
// 83-144) on top of this repo's own OBJ reader (include/xrt/obj_reader.h). tinyobjloader itself is an
// un-vendored, un-pinned dependency of the reference ("parity unpinned" for polygon triangulation
// order, see obj_reader.h). TEST INFRASTRUCTURE ONLY.
#pragma once
#include <xrt/obj_reader.h>
namespace tinyobj {
using real_t = float;
using index_t = xrt::obj::Index;
using mesh_t = xrt::obj::MeshData;
using shape_t = xrt::obj::Shape;
using material_t = xrt::obj::MaterialData;
using attrib_t = xrt::obj::Attrib;
struct ObjReaderConfig {
    bool triangulate = true;
    std::string mtl_search_path;
};
class ObjReader {
public:
    bool ParseFromFile(const std::string& filename, const ObjReaderConfig& config = ObjReaderConfig())
    {
        result_ = xrt::obj::Result();
        return xrt::obj::load(filename, config.mtl_search_path, result_);
    }
    const std::string& Error() const { return result_.error; }
    const std::string& Warning() const { return result_.warning; }
    const attrib_t& GetAttrib() const { return result_.attrib; }
    const std::vector<shape_t>& GetShapes() const { return result_.shapes; }
    const std::vector<material_t>& GetMaterials() const { return result_.materials; }
private:
    xrt::obj::Result result_;
};
} // namespace tinyobj
