"""Detector geometry (set-up time, host): arms, detector tensor, vertex.
Mirrors bilby/gw/detector/geometry.py (InterferometerGeometry), bilby/gw/geometry.py:51-115
(calculate_arm, detector_tensor) and bilby/gw/utils.py:59-88 (get_vertex_position_geocentric)."""
import numpy as np


def calculate_arm(arm_tilt, arm_azimuth, longitude, latitude):
    e_long = np.array([-np.sin(longitude), np.cos(longitude), 0.0])
    e_lat = np.array([-np.sin(latitude) * np.cos(longitude), -np.sin(latitude) * np.sin(longitude),
                      np.cos(latitude)])
    e_h = np.array([np.cos(latitude) * np.cos(longitude), np.cos(latitude) * np.sin(longitude),
                    np.sin(latitude)])
    return (np.cos(arm_tilt) * np.cos(arm_azimuth) * e_long + np.cos(arm_tilt) * np.sin(arm_azimuth) * e_lat
            + np.sin(arm_tilt) * e_h)


def get_vertex_position_geocentric(latitude, longitude, elevation):
    semi_major_axis = 6378137
    semi_minor_axis = 6356752.314
    radius = semi_major_axis ** 2 * (semi_major_axis ** 2 * np.cos(latitude) ** 2
                                     + semi_minor_axis ** 2 * np.sin(latitude) ** 2) ** (-0.5)
    x_comp = (radius + elevation) * np.cos(latitude) * np.cos(longitude)
    y_comp = (radius + elevation) * np.cos(latitude) * np.sin(longitude)
    z_comp = ((semi_minor_axis / semi_major_axis) ** 2 * radius + elevation) * np.sin(latitude)
    return np.array([x_comp, y_comp, z_comp])


class InterferometerGeometry:
    def __init__(self, length, latitude, longitude, elevation, xarm_azimuth, yarm_azimuth,
                 xarm_tilt=0., yarm_tilt=0.):
        self.length = length
        self.latitude = latitude            # degrees, like the reference's constructor
        self.longitude = longitude
        self.elevation = elevation
        self.xarm_azimuth = xarm_azimuth
        self.yarm_azimuth = yarm_azimuth
        self.xarm_tilt = xarm_tilt
        self.yarm_tilt = yarm_tilt

    @property
    def latitude_radians(self):
        return self.latitude * np.pi / 180

    @property
    def longitude_radians(self):
        return self.longitude * np.pi / 180

    @property
    def x(self):
        return calculate_arm(self.xarm_tilt, self.xarm_azimuth * np.pi / 180, self.longitude_radians,
                             self.latitude_radians)

    @property
    def y(self):
        return calculate_arm(self.yarm_tilt, self.yarm_azimuth * np.pi / 180, self.longitude_radians,
                             self.latitude_radians)

    @property
    def vertex(self):
        return get_vertex_position_geocentric(self.latitude_radians, self.longitude_radians, self.elevation)

    @property
    def detector_tensor(self):
        x, y = self.x, self.y
        return (np.outer(x, x) - np.outer(y, y)) / 2

    def unit_vector_along_arm(self, arm):
        if arm == "x":
            return self.x
        if arm == "y":
            return self.y
        raise ValueError("Arm must either be 'x' or 'y'.")
