// ref_ext.hpp — the three classes BASELINE.json's configs need but the reference does not
// contain, written ON TOP OF THE REFERENCE'S OWN HEADERS in
// the reference's conventions (double vec3, interval, hit_record, random_double()), with
// translate (src/hittable/hittable.hpp:74-117) as the style template.  Include after the
// reference headers.  Test infrastructure only (built into oracle/_ref).
#ifndef RTB200_REF_EXT_HPP
#define RTB200_REF_EXT_HPP

// Appendix B.3
class isotropic : public material
{
public:
  isotropic(const color &albedo) : tex(std::make_shared<solid_color>(albedo)) {}
  isotropic(std::shared_ptr<texture> tex) : tex(tex) {}

  bool scatter(const ray &r_in, const hit_record &rec, color &attenuation, ray &scattered) const override
  {
    scattered = ray(rec.p, random_unit_vector(), r_in.time());
    attenuation = tex->value(rec.u, rec.v, rec.p);
    return true;
  }

private:
  std::shared_ptr<texture> tex;
};

// Appendix B.1
class rotate_y : public hittable
{
public:
  rotate_y(std::shared_ptr<hittable> object, double angle) : object(object)
  {
    auto radians = degrees_to_radians(angle);
    sin_theta = std::sin(radians);
    cos_theta = std::cos(radians);
    bbox = object->bounding_box();

    point3 min(infinity, infinity, infinity);
    point3 max(-infinity, -infinity, -infinity);
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++)
        for (int k = 0; k < 2; k++)
        {
          auto x = i * bbox.x.max + (1 - i) * bbox.x.min;
          auto y = j * bbox.y.max + (1 - j) * bbox.y.min;
          auto z = k * bbox.z.max + (1 - k) * bbox.z.min;
          auto newx = cos_theta * x + sin_theta * z;
          auto newz = -sin_theta * x + cos_theta * z;
          vec3 tester(newx, y, newz);
          for (int c = 0; c < 3; c++)
          {
            min[c] = std::fmin(min[c], tester[c]);
            max[c] = std::fmax(max[c], tester[c]);
          }
        }
    bbox = aabb(min, max);
  }

  bool hit(const ray &r, interval ray_t, hit_record &rec) const override
  {
    auto origin = point3((cos_theta * r.origin().x()) - (sin_theta * r.origin().z()), r.origin().y(),
                         (sin_theta * r.origin().x()) + (cos_theta * r.origin().z()));
    auto direction = vec3((cos_theta * r.direction().x()) - (sin_theta * r.direction().z()), r.direction().y(),
                          (sin_theta * r.direction().x()) + (cos_theta * r.direction().z()));
    ray rotated_r(origin, direction, r.time());
    if (!object->hit(rotated_r, ray_t, rec))
      return false;
    rec.p = point3((cos_theta * rec.p.x()) + (sin_theta * rec.p.z()), rec.p.y(),
                   (-sin_theta * rec.p.x()) + (cos_theta * rec.p.z()));
    rec.normal = vec3((cos_theta * rec.normal.x()) + (sin_theta * rec.normal.z()), rec.normal.y(),
                      (-sin_theta * rec.normal.x()) + (cos_theta * rec.normal.z()));
    return true;
  }

  aabb bounding_box() const override { return bbox; }

private:
  std::shared_ptr<hittable> object;
  double sin_theta;
  double cos_theta;
  aabb bbox;
};

// Appendix B.2
class constant_medium : public hittable
{
public:
  constant_medium(std::shared_ptr<hittable> boundary, double density, std::shared_ptr<texture> tex)
      : boundary(boundary), neg_inv_density(-1 / density), phase_function(std::make_shared<isotropic>(tex)) {}
  constant_medium(std::shared_ptr<hittable> boundary, double density, const color &albedo)
      : boundary(boundary), neg_inv_density(-1 / density), phase_function(std::make_shared<isotropic>(albedo)) {}

  bool hit(const ray &r, interval ray_t, hit_record &rec) const override
  {
    hit_record rec1, rec2;
    if (!boundary->hit(r, interval::universe, rec1))
      return false;
    if (!boundary->hit(r, interval(rec1.t + 0.0001, infinity), rec2))
      return false;
    if (rec1.t < ray_t.min)
      rec1.t = ray_t.min;
    if (rec2.t > ray_t.max)
      rec2.t = ray_t.max;
    if (rec1.t >= rec2.t)
      return false;
    if (rec1.t < 0)
      rec1.t = 0;
    auto ray_length = r.direction().length();
    auto distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
    auto hit_distance = neg_inv_density * std::log(random_double());
    if (hit_distance > distance_inside_boundary)
      return false;
    rec.t = rec1.t + hit_distance / ray_length;
    rec.p = r.at(rec.t);
    rec.normal = vec3(1, 0, 0);
    rec.front_face = true;
    rec.mat = phase_function;
    rec.u = 0; // the book leaves u, v unset (ray_color's record is an uninitialised local): defined here as 0, 0 — the
    rec.v = 0; // convention of this repo's device code (surface_at, medium branch) and of oracle.cpp; only a uv-dependent
               // (image) phase-function texture can tell, and the reference has neither class
    return true;
  }

  aabb bounding_box() const override { return boundary->bounding_box(); }

private:
  std::shared_ptr<hittable> boundary;
  double neg_inv_density;
  std::shared_ptr<material> phase_function;
};

#endif
