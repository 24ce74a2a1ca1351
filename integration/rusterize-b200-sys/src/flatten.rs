//! `&[geo::Geometry<f64>]` -> the SoA form of `rz_geom_soa` (include/rz_b200.h), applying the pooling rules of the
//! reference's `Burn` impls (rust/src/rasterization/burn_geometry.rs:24-243):
//!
//! * Polygon / MultiPolygon / Rect / Triangle -> ONE polygon part holding every ring of every member polygon
//!   (even-odd fill over the pooled edges, :83-87, :127-133; Rect / Triangle via `to_polygon`, :50-55);
//! * LineString / MultiLineString / Line -> ONE line part, one sequence per member line (:56-59, :196-199);
//! * Point / MultiPoint -> ONE point part (:31-44);
//! * GeometryCollection -> its members' parts one after another, recursively (:64-74): members are burned
//!   independently, in order.
//!
//! Geometry `i` of the slice is geometry `i` of the SoA (a geometry without coordinates has no parts), so `field`
//! and `by` arrays keep their indices.
use crate::{RzGeomSoa, RZ_PART_LINE, RZ_PART_POINT, RZ_PART_POLYGON};
use geo_types::{Geometry, LineString, Polygon};

#[derive(Default, Debug, Clone)]
pub struct GeomSoa {
    pub geom_part_off: Vec<u64>,
    pub part_kind: Vec<u8>,
    pub part_seq_off: Vec<u64>,
    pub seq_coord_off: Vec<u64>,
    pub x: Vec<f64>,
    pub y: Vec<f64>,
}

impl GeomSoa {
    fn seq<'a>(&mut self, coords: impl IntoIterator<Item = &'a geo_types::Coord<f64>>) {
        for c in coords {
            self.x.push(c.x);
            self.y.push(c.y);
        }
        self.seq_coord_off.push(self.x.len() as u64);
    }
    fn end_part(&mut self, kind: u8) {
        self.part_kind.push(kind);
        self.part_seq_off.push((self.seq_coord_off.len() - 1) as u64);
    }
    fn rings(&mut self, p: &Polygon<f64>) {
        self.seq(p.exterior().coords()); // geo_types::Polygon::new has already closed every ring
        for hole in p.interiors() {
            self.seq(hole.coords());
        }
    }
    fn geometry(&mut self, g: &Geometry<f64>) {
        match g {
            Geometry::Point(p) => {
                self.seq(std::iter::once(&p.0));
                self.end_part(RZ_PART_POINT);
            }
            Geometry::MultiPoint(mp) => {
                self.seq(mp.iter().map(|p| &p.0));
                self.end_part(RZ_PART_POINT);
            }
            Geometry::Polygon(p) => {
                self.rings(p);
                self.end_part(RZ_PART_POLYGON);
            }
            Geometry::MultiPolygon(mp) => {
                for p in mp {
                    self.rings(p);
                }
                self.end_part(RZ_PART_POLYGON);
            }
            Geometry::Rect(r) => {
                self.rings(&r.to_polygon());
                self.end_part(RZ_PART_POLYGON);
            }
            Geometry::Triangle(t) => {
                self.rings(&t.to_polygon());
                self.end_part(RZ_PART_POLYGON);
            }
            Geometry::LineString(l) => {
                self.seq(l.coords());
                self.end_part(RZ_PART_LINE);
            }
            Geometry::MultiLineString(ml) => {
                for l in ml {
                    self.seq(l.coords());
                }
                self.end_part(RZ_PART_LINE);
            }
            Geometry::Line(l) => {
                let ls = LineString::new(vec![l.start, l.end]);
                self.seq(ls.coords());
                self.end_part(RZ_PART_LINE);
            }
            Geometry::GeometryCollection(gc) => {
                for member in gc {
                    self.geometry(member);
                }
            }
        }
    }
    /// The borrowed C view; valid while `self` is alive and unmodified.
    pub fn as_raw(&self) -> RzGeomSoa {
        RzGeomSoa {
            n_geoms: (self.geom_part_off.len() - 1) as u64,
            n_parts: self.part_kind.len() as u64,
            n_seqs: (self.seq_coord_off.len() - 1) as u64,
            n_coords: self.x.len() as u64,
            geom_part_off: self.geom_part_off.as_ptr(),
            part_kind: self.part_kind.as_ptr(),
            part_seq_off: self.part_seq_off.as_ptr(),
            seq_coord_off: self.seq_coord_off.as_ptr(),
            x: self.x.as_ptr(),
            y: self.y.as_ptr(),
        }
    }
}

pub fn flatten(geoms: &[Geometry<f64>]) -> GeomSoa {
    let mut s = GeomSoa::default();
    s.geom_part_off.reserve(geoms.len() + 1);
    s.geom_part_off.push(0);
    s.part_seq_off.push(0);
    s.seq_coord_off.push(0);
    for g in geoms {
        s.geometry(g);
        s.geom_part_off.push(s.part_kind.len() as u64);
    }
    s
}

#[cfg(test)]
mod tests {
    use super::*;
    use geo_types::{coord, Geometry, GeometryCollection, LineString, MultiPoint, Point, Polygon};

    #[test]
    fn collection_members_become_consecutive_parts() {
        let poly = Polygon::new(LineString::from(vec![(0., 0.), (4., 0.), (4., 4.)]), vec![]);
        let gc = GeometryCollection::new_from(vec![
            Geometry::Point(Point::new(1., 2.)),
            Geometry::Polygon(poly.clone()),
            Geometry::LineString(LineString::from(vec![(0., 0.), (3., 3.)])),
        ]);
        let s = flatten(&[Geometry::GeometryCollection(gc), Geometry::MultiPoint(MultiPoint::new(vec![]))]);
        assert_eq!(s.geom_part_off, vec![0, 3, 4]);
        assert_eq!(s.part_kind, vec![RZ_PART_POINT, RZ_PART_POLYGON, RZ_PART_LINE, RZ_PART_POINT]);
        assert_eq!(s.seq_coord_off, vec![0, 1, 5, 7, 7]); // the ring was closed by Polygon::new
        assert_eq!(s.x[1..5], [0., 4., 4., 0.]);
        let _ = coord! { x: 0., y: 0. };
    }
}
